"""Which ways of issuing an NCCL all-reduce survive CUDA-graph capture on this stack?  (2 ranks; diagnostic for parallel.py)
  torchrun --nproc-per-node 2 scripts/dp_capture_probe.py"""
import os
import threading

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    buf = torch.ones(1 << 20, device="cuda")
    w = torch.nn.Parameter(torch.ones(1 << 16, device="cuda"))
    side = torch.cuda.Stream()

    def on_main():
        dist.all_reduce(buf)

    def on_side():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dist.all_reduce(buf)
        cur.wait_stream(side)

    def on_side_async():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            h = dist.all_reduce(buf, async_op=True)
        h.wait()
        cur.wait_stream(side)

    def in_thread(fn):
        def run():
            cur = torch.cuda.current_stream()
            err = []

            def body():
                try:
                    torch.cuda.set_device(local)
                    with torch.cuda.stream(cur):
                        fn()
                except Exception as e:          # noqa
                    err.append(e)
            t = threading.Thread(target=body); t.start(); t.join()
            if err:
                raise err[0]
        return run

    class Hook(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, fn):
            ctx.fn = fn
            return x * 2

        @staticmethod
        def backward(ctx, g):
            ctx.fn()
            return g * 2, None

    def in_backward(fn):
        def run():
            Hook.apply(w, fn).sum().backward()
        return run

    cases = [("main stream, main thread", on_main), ("side stream, main thread", on_side), ("side stream async_op, main thread", on_side_async),
             ("main stream, other thread", in_thread(on_main)), ("side stream, other thread", in_thread(on_side)),
             ("main stream, autograd backward", in_backward(on_main)), ("side stream, autograd backward", in_backward(on_side))]
    for mode in ("global", "thread_local", "relaxed"):
        for name, fn in cases:
            gs = torch.cuda.Stream()
            ok, msg = True, ""
            try:
                with torch.cuda.stream(gs):
                    fn()                              # eager warm-up on the capture stream
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                buf.fill_(1.0)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=gs, capture_error_mode=mode):
                    fn()
                g.replay(); g.replay()
                torch.cuda.synchronize()
                ok = abs(float(buf[0]) - world ** 2) < 1e-3
                msg = f"value {float(buf[0])}"
                g.reset()
            except Exception as e:                    # noqa
                ok, msg = False, f"{type(e).__name__}: {str(e).splitlines()[0][:120]}"
                try:
                    torch.cuda.synchronize()
                except Exception:                     # noqa
                    pass
            if rank == 0:
                print(f"[{mode:12s}] {name:36s} {'OK ' if ok else 'FAIL'} {msg}", flush=True)
            dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
