// Launch lists: one C call issues the launches of a residual block's forward or backward pass.
//
// The step is ~560 launches of this library; from Python every launch costs a ctypes call plus the interpreter's walk to it.
// us3d_run_ops takes the already-resolved arguments of a sequence of entry points (device pointers, sizes) and calls them in
// order on one stream — same kernels, same arguments, same order as the call-by-call route, so results are bit-identical.
// The host-side mirror is unscene3d_b200/engine/blocks.py (the reference's BasicBlock, models/modules/resnet_block.py:24-64).
#include "common.cuh"

using namespace us3d;

extern "C" int us3d_run_ops(const us3d_op_t *ops, int n_ops, void *stream) {
    US3D_CHECK_ARG(n_ops >= 0 && (n_ops == 0 || ops != nullptr), "run_ops: bad list");
    for (int i = 0; i < n_ops; ++i) {
        const us3d_op_t &o = ops[i];
        const void *const *p = o.p;
        const long long *v = o.v;
        int rc = 0;
        switch (o.kind) {
            case US3D_OP_CONV: {
                // p: x_hi, x_lo, nbr, wpack, bias, out_rows, y, tile_mask, partition, workspace,
                //    [10] bn ws, mean, invstd, running_mean, running_var, num_batches_tracked (bn ws == NULL: no statistics)
                // v: n_in, n_rows, kvol, cin, cout, passes, ldy, accumulate, workspace_bytes;   f: eps, momentum
                us3d_bn_fuse_t bn;
                bn.ws = (double *)p[10]; bn.mean = (float *)p[11]; bn.invstd = (float *)p[12];
                bn.running_mean = (float *)p[13]; bn.running_var = (float *)p[14]; bn.num_batches_tracked = (long long *)p[15];
                bn.eps = o.f[0]; bn.momentum = o.f[1];
                rc = us3d_spconv_gather_mt_bn(p[0], p[1], (int)v[0], (const int32_t *)p[2], (int)v[1], (int)v[2], p[3], (int)v[3], (int)v[4],
                                              (int)v[5], (const float *)p[4], (const int32_t *)p[5], (float *)p[6], (int)v[6], (int)v[7],
                                              (const uint32_t *)p[7], (const int32_t *)p[8], (void *)p[9], v[8],
                                              p[10] != nullptr ? &bn : nullptr, stream);
                break;
            }
            case US3D_OP_BN_APPLY:
                // p: x, mean, invstd, gamma, beta, residual, y, hi, lo;   v: ldx, n, c, ldr, relu, ldy
                rc = us3d_bn_apply_planes((const float *)p[0], (int)v[0], (int)v[1], (int)v[2], (const float *)p[1], (const float *)p[2],
                                          (const float *)p[3], (const float *)p[4], (const float *)p[5], (int)v[3], (int)v[4], (float *)p[6],
                                          (int)v[5], (void *)p[7], (void *)p[8], stream);
                break;
            case US3D_OP_BN_BACKWARD:
                // p: dy, x, y, mean, invstd, gamma, ws, dx, dres, dgamma, dbeta, dx_hi, dx_lo
                // v: lddy, ldx, ldy, n, c, relu, batch_terms, lddx, lddres
                rc = us3d_bn_backward_planes((const float *)p[0], (int)v[0], (const float *)p[1], (int)v[1], (const float *)p[2], (int)v[2],
                                             (int)v[3], (int)v[4], (const float *)p[3], (const float *)p[4], (const float *)p[5], (int)v[5],
                                             (int)v[6], (double *)p[6], (float *)p[7], (int)v[7], (float *)p[8], (int)v[8], (float *)p[9],
                                             (float *)p[10], (void *)p[11], (void *)p[12], stream);
                break;
            case US3D_OP_WGRAD:
                // p: x_hi, x_lo, dy_hi, dy_lo, nbr, dw, tile_mask, dy_rows;   v: n_rows, kvol, cin, cout, passes
                rc = us3d_spconv_wgrad_planes(p[0], p[1], p[2], p[3], (const int32_t *)p[4], (int)v[0], (int)v[1], (float *)p[5], (int)v[2],
                                              (int)v[3], (int)v[4], (const uint32_t *)p[6], (const int32_t *)p[7], stream);
                break;
            case US3D_OP_ADD:
                // p: a, b, z;   v: numel
                rc = us3d_add((const float *)p[0], (const float *)p[1], (float *)p[2], v[0], stream);
                break;
            default:
                US3D_CHECK_ARG(false, "run_ops: unknown op kind %d at position %d", o.kind, i);
        }
        if (rc != 0) return rc;
    }
    return 0;
}

// The same list as flat arrays (what a ctypes caller can fill with one conversion): a = n_ops x [kind, p[16], v[10]] as int64,
// f = n_ops x [f0, f1] as double.
extern "C" int us3d_run_ops_flat(const long long *a, const double *f, int n_ops, void *stream) {
    US3D_CHECK_ARG(n_ops >= 0 && (n_ops == 0 || (a != nullptr && f != nullptr)), "run_ops_flat: bad list");
    for (int i = 0; i < n_ops; ++i) {
        const long long *r = a + (size_t)i * 27;
        us3d_op_t o;
        o.kind = (int)r[0];
        for (int k = 0; k < 16; ++k) o.p[k] = reinterpret_cast<const void *>(static_cast<uintptr_t>(r[1 + k]));
        for (int k = 0; k < 10; ++k) o.v[k] = r[17 + k];
        o.f[0] = (float)f[2 * i];
        o.f[1] = (float)f[2 * i + 1];
        const int rc = us3d_run_ops(&o, 1, stream);
        if (rc != 0) return rc;
    }
    return 0;
}
