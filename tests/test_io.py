"""On-disk formats either side of the path (dsr_b200.io, csrc/resize.cu) against the numpy restatement of the reference's
conversions (oracle/ref_io.py).  Integer outputs must be BIT-EXACT, float outputs bit-exact too (same IEEE operations)."""
import numpy as np
import pytest
import torch

from oracle import ref_io, ref_step


def test_png_u16_roundtrip_host(tmp_path):
    from dsr_b200 import io
    a = (np.random.RandomState(0).rand(48, 64) * 65535).astype(np.uint16)
    f = str(tmp_path / "d.png")
    io.write_png_u16(f, a)
    assert np.array_equal(io.read_png_u16(f), a)


@pytest.mark.gpu
def test_input_conversions_bit_exact(built_lib):
    from dsr_b200 import io
    rs = np.random.RandomState(1)
    d = rs.randint(0, 9000, size=(3, 37, 53)).astype(np.uint16)
    d[0, :3] = 0; d[1, 5] = 5100; d[2, 7] = 5101; d[2, 8] = 65535
    out = io.depth_from_u16(d, "cuda")
    assert out.shape == (3, 1, 37, 53)
    assert np.array_equal(out.cpu().numpy()[:, 0], ref_io.depth_from_u16(d))
    img = rs.randint(0, 256, size=(2, 19, 23, 3)).astype(np.uint8)
    assert np.array_equal(io.image_from_u8(img, "cuda").cpu().numpy(), ref_io.image_from_u8(img))


@pytest.mark.gpu
def test_export_bit_exact_and_save_all(built_lib, tmp_path):
    from dsr_b200 import io
    g = torch.Generator().manual_seed(2)
    pred = torch.rand(2, 1, 64, 40, generator=g) * 2.4 - 1.2           # values outside [-1, 1] exercise the clip
    pred[0, 0, 20, :5] = torch.tensor([-1.0, 1.0, 0.0, 0.99999994, -0.99999994])
    got = io.depth_to_u16(pred.cuda(), 16)
    assert got.dtype == np.uint16 and got.shape == (2, 32, 40)
    assert np.array_equal(got, ref_io.depth_to_u16(pred.numpy(), 16))
    files = io.save_predictions(pred.cuda(), ["/data/x/scene0001_00.png", "b/frame.7.jpg"], str(tmp_path) + "/", 16)
    assert [f.split("/")[-1] for f in files] == ["scene0001_00.png", "frame.png"]          # main_model.py:329-330 naming
    assert np.array_equal(io.read_png_u16(files[0]), got[0]) and np.array_equal(io.read_png_u16(files[1]), got[1])


@pytest.mark.gpu
def test_model_save_all_writes_pngs(built_lib, tmp_path):
    """--save_all in the test stage (main_model.py:321-333) through the public API"""
    from util import build_host_model, rehome
    host = build_host_model(1, 128, 128, save_all=True, save_image_folder=str(tmp_path) + "/")
    model = rehome(host, host.opt, [0])
    model.eval()
    batch = ref_step.synthetic_batch(1, 128, 128, seed=3, depth_kind="smooth")
    batch["B_paths"] = ["/somewhere/real_0042.png"]
    with torch.no_grad():
        model.set_input(batch)
        model.forward("test")
    out = __import__("dsr_b200.io", fromlist=["io"]).read_png_u16(str(tmp_path / "real_0042.png"))
    assert out.shape == (96, 128) and out.dtype == np.uint16
    assert np.array_equal(out, ref_io.depth_to_u16(model.pred_real_depth.detach().cpu().numpy(), 16)[0])
