// Displacement-field self/cross warps: one scaling-and-squaring step, compose, VecInt driver.
//   out = scale*own + interp(scale*src, p + scale*own)
// Direct-gather formulation: the source field (7.4 MB at 80x80x96, 59 MB at 160x160x192) is
// L2-resident and neighbouring lanes share 128-B lines, so the 24 gathers per voxel are
// L1/L2 hits; HBM sees one read of `own` and one write of `out` (24 B/voxel).
#include "dfm_common.cuh"

namespace dfm {

template <int VEC, int INTERP, bool IN_CL, bool OUT_CL>
__global__ void __launch_bounds__(256)
k_field_warp_add(const float *__restrict__ src, const float *__restrict__ own, float *__restrict__ out,
                 int Xs, int Ys, int Zs, int X, int Y, int Z, float scale, FastDiv zvdiv,
                 uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zvdiv);
    const uint32_t z = (p - y * zvdiv.d) * VEC;
    const uint32_t x = blockIdx.y;
    const size_t N = (size_t)X * Y * Z, Ns = (size_t)Xs * Ys * Zs;
    const size_t vox = ((size_t)x * Y + y) * Z + z;
    const float *ownb = own + (size_t)blockIdx.z * 3 * N;
    const float *srcb = src + (size_t)blockIdx.z * 3 * Ns;
    float *outb = out + (size_t)blockIdx.z * 3 * N;

    float v[3][VEC];
    if (IN_CL) {
        float a[3 * VEC];
        if (VEC == 4) {
            const float4 *q = reinterpret_cast<const float4 *>(ownb + vox * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 t = __ldg(q + k);
                a[4 * k] = t.x; a[4 * k + 1] = t.y; a[4 * k + 2] = t.z; a[4 * k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] = __ldg(ownb + vox * 3 + k);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c][i] = a[i * 3 + c];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (VEC == 4) {
                float4 t = __ldg(reinterpret_cast<const float4 *>(ownb + c * N + vox));
                v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
            } else {
                v[c][0] = __ldg(ownb + c * N + vox);
            }
        }
    }

    float r[3][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float v0 = __fmul_rn(scale, v[0][i]);
        const float v1 = __fmul_rn(scale, v[1][i]);
        const float v2 = __fmul_rn(scale, v[2][i]);
        const float lx = __fadd_rn((float)x, v0);
        const float ly = __fadd_rn((float)y, v1);
        const float lz = __fadd_rn((float)(z + i), v2);
        float acc[3];
        if (INTERP == DFM_LINEAR) {
            uint32_t off[8];
            float w[8];
            tri_setup(lx, ly, lz, Xs, Ys, Zs, off, w);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float val[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    val[k] = IN_CL ? __ldg(srcb + (size_t)off[k] * 3 + c) : __ldg(srcb + c * Ns + off[k]);
                acc[c] = tri_accumulate(w, val);
            }
        } else {
            const uint32_t o = ((uint32_t)axis_nearest(lx, Xs - 1) * Ys + axis_nearest(ly, Ys - 1)) * Zs +
                               axis_nearest(lz, Zs - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                acc[c] = IN_CL ? __ldg(srcb + (size_t)o * 3 + c) : __ldg(srcb + c * Ns + o);
        }
        r[0][i] = __fadd_rn(v0, __fmul_rn(scale, acc[0]));
        r[1][i] = __fadd_rn(v1, __fmul_rn(scale, acc[1]));
        r[2][i] = __fadd_rn(v2, __fmul_rn(scale, acc[2]));
    }

    if (OUT_CL) {
        if (VEC == 4) {
            float a[12];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) a[i * 3 + c] = r[c][i];
            float4 *q = reinterpret_cast<float4 *>(outb + vox * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) outb[vox * 3 + c] = r[c][0];
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (VEC == 4)
                *reinterpret_cast<float4 *>(outb + c * N + vox) = make_float4(r[c][0], r[c][1], r[c][2], r[c][3]);
            else
                outb[c * N + vox] = r[c][0];
        }
    }
}

template <int VEC, int INTERP>
static int launch_fwa(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs,
                      int X, int Y, int Z, float scale, unsigned flags, cudaStream_t st) {
    const uint32_t zv = Z / VEC;
    const uint32_t plane = (uint32_t)Y * zv;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(zv);
    const bool icl = flags & DFM_FIELD_IN_CL, ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(I, O) k_field_warp_add<VEC, INTERP, I, O><<<grid, block, 0, st>>>( \
        src, own, out, Xs, Ys, Zs, X, Y, Z, scale, fd, plane)
    if (icl) { if (ocl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (ocl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_field_warp_add");
}

static int validate_grid(const char *who, int B, int X, int Y, int Z) {
    DFM_REQUIRE(B >= 0 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL, "%s: bad shape B=%d X=%d Y=%d Z=%d", who, B, X, Y, Z);
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "%s: B and X must be <= 65535", who);
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31), DFM_EINVAL, "%s: volume too large (>= 2^31 voxels)", who);
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "%s: Y*Z*Z must be < 2^32", who);
    return DFM_OK;
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_field_warp_add(const float *src, const float *own, float *out, int B, int Xs, int Ys,
                                  int Zs, int X, int Y, int Z, float scale, int interp, unsigned flags,
                                  void *stream) {
    int rc;
    if ((rc = validate_grid("dfm_field_warp_add", B, X, Y, Z))) return rc;
    if ((rc = validate_grid("dfm_field_warp_add(src)", B, Xs, Ys, Zs))) return rc;
    DFM_REQUIRE(src && own && out, DFM_EINVAL, "dfm_field_warp_add: null pointer");
    DFM_REQUIRE(interp == DFM_LINEAR || interp == DFM_NEAREST, DFM_EINVAL, "dfm_field_warp_add: interp %d", interp);
    DFM_REQUIRE(out != own && out != src, DFM_EINVAL, "dfm_field_warp_add: out must not alias an input");
    int ex;
    DFM_REQUIRE(frexpf(scale, &ex) == 0.5f, DFM_EINVAL, "dfm_field_warp_add: scale %g is not a power of two", scale);
    if (B == 0) return DFM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (Z % 4 == 0) && aligned16(own) && aligned16(out);
    if (interp == DFM_LINEAR)
        return vec4 ? launch_fwa<4, DFM_LINEAR>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st)
                    : launch_fwa<1, DFM_LINEAR>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st);
    return vec4 ? launch_fwa<4, DFM_NEAREST>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st)
                : launch_fwa<1, DFM_NEAREST>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st);
}

extern "C" size_t dfm_vecint_workspace_bytes(int B, int X, int Y, int Z, int nsteps, int save_steps) {
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0 || nsteps <= 0) return 0;
    const size_t one = (size_t)B * 3 * X * Y * Z * sizeof(float);
    if (save_steps) return one * (size_t)nsteps;
    return nsteps >= 2 ? one : 0;
}

extern "C" int dfm_vecint_fwd(const float *svf, float *out, float *work, int B, int X, int Y, int Z,
                              int nsteps, int save_steps, unsigned flags, void *stream) {
    int rc;
    if ((rc = validate_grid("dfm_vecint_fwd", B, X, Y, Z))) return rc;
    DFM_REQUIRE(nsteps >= 0 && nsteps <= 30, DFM_EINVAL, "dfm_vecint_fwd: nsteps %d out of range [0, 30]", nsteps);
    DFM_REQUIRE(svf && out, DFM_EINVAL, "dfm_vecint_fwd: null pointer");
    DFM_REQUIRE(svf != out, DFM_EINVAL, "dfm_vecint_fwd: out must not alias svf");
    if (B == 0) return DFM_OK;
    const size_t n = (size_t)B * 3 * X * Y * Z;
    const unsigned in_cl = flags & DFM_FIELD_IN_CL, out_cl = flags & DFM_FIELD_OUT_CL;
    if (nsteps == 0) {   // integrate_vec with nb_steps = 0 is the identity (v / 2**0)
        if (in_cl == (out_cl ? DFM_FIELD_IN_CL : 0u)) {
            cudaError_t e = cudaMemcpyAsync(out, svf, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
            DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_vecint_fwd: %s", cudaGetErrorString(e));
            return DFM_OK;
        }
        return in_cl ? dfm_cl_to_planar(svf, out, B, 3, (size_t)X * Y * Z, 4, stream)
                     : dfm_planar_to_cl(svf, out, B, 3, (size_t)X * Y * Z, 4, stream);
    }
    const size_t need = dfm_vecint_workspace_bytes(B, X, Y, Z, nsteps, save_steps);
    DFM_REQUIRE(need == 0 || work, DFM_EINVAL, "dfm_vecint_fwd: workspace of %zu bytes required", need);
    const float scale0 = ldexpf(1.f, -nsteps);
    if (save_steps) {
        // work[k] = v_k (input of step k), k = 0..nsteps-1; v_0 = svf * 2^-nsteps (planar)
        rc = scale_copy_to_planar(svf, work, B, (size_t)X * Y * Z, scale0, in_cl != 0, (cudaStream_t)stream);
        if (rc) return rc;
        for (int k = 0; k < nsteps; ++k) {
            const float *vin = work + (size_t)k * n;
            float *vout = (k == nsteps - 1) ? out : work + (size_t)(k + 1) * n;
            unsigned f = (k == nsteps - 1) ? out_cl : 0u;
            rc = dfm_field_warp_add(vin, vin, vout, B, X, Y, Z, X, Y, Z, 1.f, DFM_LINEAR, f, stream);
            if (rc) return rc;
        }
        return DFM_OK;
    }
    // ping-pong between `work` and `out` so that the last step lands in `out`
    const float *cur = svf;
    unsigned cur_cl = in_cl;
    for (int k = 0; k < nsteps; ++k) {
        const bool last = (k == nsteps - 1);
        float *dst = ((nsteps - 1 - k) % 2 == 0) ? out : work;
        unsigned f = cur_cl | (last ? out_cl : 0u);
        // intermediate buffers are planar; if `out` is used as an intermediate it holds planar data
        rc = dfm_field_warp_add(cur, cur, dst, B, X, Y, Z, X, Y, Z, k == 0 ? scale0 : 1.f, DFM_LINEAR, f, stream);
        if (rc) return rc;
        cur = dst;
        cur_cl = 0u;
    }
    return DFM_OK;
}
