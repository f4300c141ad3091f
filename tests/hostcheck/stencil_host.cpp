// Host build of the per-pixel stencil math (csrc/stencil_math.cuh) so the analytic forward/backward
// formulas can be checked against the oracle on a machine with no GPU.  Test infrastructure only:
// the product never calls this; the CUDA kernels include the very same header.
#include "../../depth-enhancement-and-super-resolution_b200/csrc/stencil_math.cuh"

extern "C" {
void host_normals_old_fwd(const float* d, int B, int H, int W, float scale, float* out) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float n[3];
                old_normal_fwd(d + b * plane, H, W, i, j, scale, n);
                for (int c = 0; c < 3; ++c) out[(b * 3 + c) * plane + (long)i * W + j] = n[c];
            }
}
void host_normals_old_bwd(const float* d, const float* g, int B, int H, int W, float scale, float* gd) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j)
                gd[b * plane + (long)i * W + j] = old_normal_bwd(d + b * plane, g + b * 3 * plane, plane, H, W, i, j, scale);
}
void host_normals_new_fwd(const float* d, const double* cams, int B, int H, int W, float* out) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float n[3];
                new_normal_fwd(d + b * plane, cams + b * DSR_CAM_DOUBLES, H, W, i, j, n);
                for (int c = 0; c < 3; ++c) out[(b * 3 + c) * plane + (long)i * W + j] = n[c];
            }
}
void host_normals_new_bwd(const float* d, const float* g, const double* cams, int B, int H, int W, float* gd) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j)
                gd[b * plane + (long)i * W + j] =
                    new_normal_bwd(d + b * plane, g + b * 3 * plane, plane, cams + b * DSR_CAM_DOUBLES, H, W, i, j);
}
// closed-form fp32 path for affine (pin-hole) cameras; returns 0 when a camera of the batch is not affine
int host_normals_aff_fwd(const float* d, const double* cams, int B, int H, int W, float* out) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b) {
        if (!cam_is_affine(cams + b * DSR_CAM_DOUBLES)) return 0;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float n[3];
                aff_normal_fwd(d + b * plane, cams + b * DSR_CAM_DOUBLES, H, W, i, j, n);
                for (int c = 0; c < 3; ++c) out[(b * 3 + c) * plane + (long)i * W + j] = n[c];
            }
    }
    return 1;
}
int host_normals_aff_bwd(const float* d, const float* g, const double* cams, int B, int H, int W, float* gd) {
    long plane = (long)H * W;
    for (int b = 0; b < B; ++b) {
        if (!cam_is_affine(cams + b * DSR_CAM_DOUBLES)) return 0;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j)
                gd[b * plane + (long)i * W + j] =
                    aff_normal_bwd(d + b * plane, g + b * 3 * plane, plane, cams + b * DSR_CAM_DOUBLES, H, W, i, j);
    }
    return 1;
}
void host_bilin_ac(int o, int n_out, int n_in, int* i0, int* i1, float* l0, float* l1) { bilin_ac(o, n_out, n_in, *i0, *i1, *l0, *l1); }
}
