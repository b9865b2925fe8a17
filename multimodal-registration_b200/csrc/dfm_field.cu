// Displacement-field self/cross warps: one scaling-and-squaring step, compose, VecInt driver.
//   out = scale*own + interp(scale*src, p + scale*own)
//
// Two implementations:
//  * k_field_warp_add  -- direct gathers through L1/L2 (any layout, any shape, nearest/linear).
//    Lanes of a warp own 32 consecutive z of one row, a thread owns ROWS consecutive y rows,
//    so every gather instruction touches 1-2 lines and the ROWS chains are independent (ILP).
//  * k_ss_brick (dfm_brick.cu) -- planar linear path: the bounding box of the tile's sample
//    locations is staged in shared memory by one TMA box load; see there.
#include <stdlib.h>

#include "dfm_common.cuh"

namespace dfm {

template <int ROWS, int INTERP, bool IN_CL, bool OUT_CL>
__global__ void __launch_bounds__(256, 4)
k_field_warp_add(const float *__restrict__ src, const float *__restrict__ own, float *__restrict__ out,
                 int Xs, int Ys, int Zs, int X, int Y, int Z, float scale, FastDiv zdiv,
                 uint32_t plane_items, float *absmax) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t yy = p < plane_items ? fast_div(p, zdiv) : 0xffffffffu / ROWS;      // past the end: no rows
    float am = 0.f;                                                  // max |out| (absmax != nullptr)
    const uint32_t z = p - yy * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ns = (uint32_t)Xs * Ys * Zs;
    const float *ownb = own + (size_t)blockIdx.z * 3 * N;
    const float *srcb = src + (size_t)blockIdx.z * 3 * Ns;
    float *outb = out + (size_t)blockIdx.z * 3 * N;
    const float *s0 = srcb, *s1 = IN_CL ? srcb + 1 : srcb + Ns, *s2 = IN_CL ? srcb + 2 : srcb + 2 * (size_t)Ns;
    const float fx = (float)x, fz = (float)z;
    const bool fast = Xs >= 2 && Ys >= 2 && Zs >= 2;                 // uniform
    const uint32_t gy = (IN_CL ? 3u : 1u) * Zs, gx = gy * Ys;      // corner strides (elements)
    const size_t cs = IN_CL ? 1 : Ns;                                // component stride

#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t y = yy * ROWS + r;
        if (y >= (uint32_t)Y) break;
        const uint32_t vox = (x * Y + y) * Z + z;
        float v0, v1, v2;
        if (IN_CL) {
            v0 = __ldg(ownb + (size_t)vox * 3); v1 = __ldg(ownb + (size_t)vox * 3 + 1); v2 = __ldg(ownb + (size_t)vox * 3 + 2);
        } else {
            v0 = __ldg(ownb + vox); v1 = __ldg(ownb + N + vox); v2 = __ldg(ownb + 2 * (size_t)N + vox);
        }
        v0 = __fmul_rn(scale, v0); v1 = __fmul_rn(scale, v1); v2 = __fmul_rn(scale, v2);
        const float lx = __fadd_rn(fx, v0), ly = __fadd_rn((float)y, v1), lz = __fadd_rn(fz, v2);
        float a0, a1, a2;
        if (INTERP == DFM_LINEAR) {
            float w[8], val[8];
            if (fast) {
                const uint32_t base = tri_setup_fast(lx, ly, lz, Xs, Ys, Zs, w);
                const uint32_t es = IN_CL ? 3u : 1u;
                const float *g = s0 + (size_t)base * es;
                gather8(g, gy, gx, es, val);
                a0 = tri_accumulate(w, val);
                gather8(g + cs, gy, gx, es, val);
                a1 = tri_accumulate(w, val);
                gather8(g + 2 * cs, gy, gx, es, val);
                a2 = tri_accumulate(w, val);
            } else {
                uint32_t off[8];
                tri_setup(lx, ly, lz, Xs, Ys, Zs, off, w);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = __ldg(s0 + (IN_CL ? off[k] * 3 : off[k]));
                a0 = tri_accumulate(w, val);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = __ldg(s1 + (IN_CL ? off[k] * 3 : off[k]));
                a1 = tri_accumulate(w, val);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = __ldg(s2 + (IN_CL ? off[k] * 3 : off[k]));
                a2 = tri_accumulate(w, val);
            }
        } else {
            uint32_t o = ((uint32_t)axis_nearest(lx, Xs - 1) * Ys + axis_nearest(ly, Ys - 1)) * Zs + axis_nearest(lz, Zs - 1);
            if (IN_CL) o *= 3;
            a0 = __ldg(s0 + o); a1 = __ldg(s1 + o); a2 = __ldg(s2 + o);
        }
        const float r0 = __fadd_rn(v0, __fmul_rn(scale, a0));
        const float r1 = __fadd_rn(v1, __fmul_rn(scale, a1));
        const float r2 = __fadd_rn(v2, __fmul_rn(scale, a2));
        am = absmax_fold(absmax_fold(absmax_fold(am, r0), r1), r2);
        if (OUT_CL) {
            outb[(size_t)vox * 3] = r0; outb[(size_t)vox * 3 + 1] = r1; outb[(size_t)vox * 3 + 2] = r2;
        } else {
            outb[vox] = r0; outb[N + vox] = r1; outb[2 * (size_t)N + vox] = r2;
        }
    }
    if (absmax) block_absmax_commit(am, absmax + blockIdx.z);        // uniform branch
}

// planar linear variant written for memory-level parallelism (see k_warp_linear1): all 24 corner
// gathers of a voxel are issued before the first use, at a register budget that keeps 5 CTAs/SM.
template <bool SCALED, bool OUT_CL>
__global__ void __launch_bounds__(256, 5)
k_field_warp_add1(const float *__restrict__ src, const float *__restrict__ own, float *__restrict__ out,
                  int Xs, int Ys, int Zs, int X, int Y, int Z, float scale, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zdiv);
    const uint32_t z = p - y * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ns = (uint32_t)Xs * Ys * Zs;
    const float *ownb = own + (size_t)blockIdx.z * 3 * N;
    const float *srcb = src + (size_t)blockIdx.z * 3 * Ns;
    float *outb = out + (size_t)blockIdx.z * 3 * N;
    const uint32_t vox = (x * Y + y) * Z + z;
    const int mxi = Xs - 1, myi = Ys - 1, mzi = Zs - 1;
    float v0 = __ldg(ownb + vox), v1 = __ldg(ownb + N + vox), v2 = __ldg(ownb + 2 * (size_t)N + vox);
    if (SCALED) { v0 = __fmul_rn(scale, v0); v1 = __fmul_rn(scale, v1); v2 = __fmul_rn(scale, v2); }
    const float lx = __fadd_rn((float)x, v0), ly = __fadd_rn((float)y, v1), lz = __fadd_rn((float)z, v2);
    const AxisF ax = axis_fast(lx, (float)mxi, mxi), ay = axis_fast(ly, (float)myi, myi), az = axis_fast(lz, (float)mzi, mzi);
    const uint32_t gy = (uint32_t)Zs, gx = (uint32_t)Ys * Zs;
    const float *g = srcb + (((uint32_t)(ax.i1 - 1) * Ys + (uint32_t)(ay.i1 - 1)) * Zs + (uint32_t)(az.i1 - 1));
    float a[3][8];
    gather8(g, gy, gx, 1u, a[0]);
    gather8(g + Ns, gy, gx, 1u, a[1]);
    gather8(g + 2 * (size_t)Ns, gy, gx, 1u, a[2]);
    float w[8];
    tri_weights(ax, ay, az, w);
    float r0 = tri_accumulate(w, a[0]), r1 = tri_accumulate(w, a[1]), r2 = tri_accumulate(w, a[2]);
    if (SCALED) { r0 = __fmul_rn(scale, r0); r1 = __fmul_rn(scale, r1); r2 = __fmul_rn(scale, r2); }
    r0 = __fadd_rn(v0, r0); r1 = __fadd_rn(v1, r1); r2 = __fadd_rn(v2, r2);
    if (OUT_CL) {
        outb[(size_t)vox * 3] = r0; outb[(size_t)vox * 3 + 1] = r1; outb[(size_t)vox * 3 + 2] = r2;
    } else {
        outb[vox] = r0; outb[N + vox] = r1; outb[2 * (size_t)N + vox] = r2;
    }
}

static int launch_fwa1(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X, int Y,
                       int Z, float scale, unsigned flags, cudaStream_t st) {
    const uint32_t plane = (uint32_t)Y * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(Z);
    const bool ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(S, O) k_field_warp_add1<S, O><<<grid, block, 0, st>>>(src, own, out, Xs, Ys, Zs, X, Y, Z, scale, fd, plane)
    if (scale == 1.f) { if (ocl) DFM_GO(false, true); else DFM_GO(false, false); }
    else              { if (ocl) DFM_GO(true, true); else DFM_GO(true, false); }
#undef DFM_GO
    return check_launch("dfm_field_warp_add(direct, mlp)");
}

template <int INTERP>
static int launch_fwa(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs,
                      int X, int Y, int Z, float scale, unsigned flags, cudaStream_t st, float *absmax = nullptr) {
    constexpr int ROWS = 4;
    const uint32_t plane = (uint32_t)((Y + ROWS - 1) / ROWS) * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(Z);
    const bool icl = flags & DFM_FIELD_IN_CL, ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(I, O) k_field_warp_add<ROWS, INTERP, I, O><<<grid, block, 0, st>>>( \
        src, own, out, Xs, Ys, Zs, X, Y, Z, scale, fd, plane, absmax)
    if (icl) { if (ocl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (ocl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_field_warp_add");
}

int validate_grid(const char *who, int B, int X, int Y, int Z) {
    DFM_REQUIRE(B >= 0 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL, "%s: bad shape B=%d X=%d Y=%d Z=%d", who, B, X, Y, Z);
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "%s: B and X must be <= 65535", who);
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 30), DFM_EINVAL, "%s: volume too large (>= 2^30 voxels)", who);
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
    if (interp == DFM_LINEAR) {
        if (brick_eligible(src, own, out, Xs, Ys, Zs, X, Y, Z, flags)) {
            // a stand-alone call (compose, a single SS step) sees full-size displacements: larger brick
            rc = launch_ss_brick(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, /*large_box=*/1, nullptr, 0.f, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
        }
        if (!(flags & DFM_FIELD_IN_CL) && Xs >= 2 && Ys >= 2 && Zs >= 2)
            return launch_fwa1(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st);
        return launch_fwa<DFM_LINEAR>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st);
    }
    return launch_fwa<DFM_NEAREST>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, flags, st);
}

extern "C" size_t dfm_vecint_workspace_bytes(int B, int X, int Y, int Z, int nsteps, int save_steps) {
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0 || nsteps <= 0) return 0;
    const size_t one = (size_t)B * 3 * X * Y * Z * sizeof(float);
    // per-item displacement maxima: of the first step's output (static halo, adjoint kernel selection) and of the outputs of the
    // two steps before the last (halo selection of the last two steps)
    const size_t bound = 3 * (((size_t)B * sizeof(float) + 255) / 256 * 256);
    if (save_steps) return one * (size_t)nsteps + bound;
    return nsteps >= 2 ? one + bound : 0;
}

// one SS step with the implementation choice: late steps (large displacements, stronger local
// deformation) get the larger brick
// bound/bscale: see launch_ss_brick; absmax (nullable): receives max |vout| per item when the step runs on the
// gather kernel (channels-last input -- the first step of an inference call)
static int ss_step(const float *vin, float *vout, int B, int X, int Y, int Z, float scale, unsigned flags,
                   int steps_left, const float *bound, float bscale, float *absmax, cudaStream_t st) {
    static const bool prefer_direct = getenv("DFM_SS_DIRECT") != nullptr;      // tuning aid
    if (prefer_direct && !(flags & DFM_FIELD_IN_CL) && X >= 2 && Y >= 2 && Z >= 2)
        return launch_fwa1(vin, vin, vout, B, X, Y, Z, X, Y, Z, scale, flags, st);
    if (brick_eligible(vin, vin, vout, X, Y, Z, X, Y, Z, flags)) {
        int rc = launch_ss_brick(vin, vin, vout, B, X, Y, Z, X, Y, Z, scale, steps_left < 3 ? 1 : 0, bound, bscale, st);
        if (rc != DFM_EUNSUPPORTED) return rc;
    }
    if (!(flags & DFM_FIELD_IN_CL) && X >= 2 && Y >= 2 && Z >= 2)
        return launch_fwa1(vin, vin, vout, B, X, Y, Z, X, Y, Z, scale, flags, st);
    return launch_fwa<DFM_LINEAR>(vin, vin, vout, B, X, Y, Z, X, Y, Z, scale, flags, st, absmax);
}

// One SS step on the plane-marching kernel (dfm_ss_march.cu).  `bound` (nullable) * bscale bounds the
// displacements of the step's input: the last two steps pick the halo-2 or the halo-3 ring PER ITEM on the
// device (both variants are launched; the CTAs of the one not selected skip the item), earlier steps always
// run halo 2 (|v| halves with every step back; outliers gather from global memory, so this is only tuning).
// `meas_in` (nullable, device): max |vin| per item as measured by the previous step -- when given it replaces the doubled bound
// for the selection; `meas_out` (nullable): this step records max |vout| per item there (zeroed by the caller).
static int march_step(const float *vin, float *vout, int B, int X, int Y, int Z, float scale, bool in_cl, bool first,
                      float *absmax, int steps_left, const float *bound, float bscale, const float *meas_in, float *meas_out,
                      cudaStream_t st) {
    // an item runs on the halo-2 ring iff its displacement figure is below the threshold: 3.2 on the doubled bound (loose by
    // construction), DFM_MARCH_THR_MEAS on the measured maximum
    static const float thr = getenv("DFM_MARCH_THR") ? (float)atof(getenv("DFM_MARCH_THR")) : 3.2f;    // tuning aid
    static const float thr_m = getenv("DFM_MARCH_THR_MEAS") ? (float)atof(getenv("DFM_MARCH_THR_MEAS")) : 3.0f;
    float *mo = first ? absmax : meas_out;
    if (steps_left >= 2 || !(bound || meas_in))
        return launch_ss_march(vin, vout, B, X, Y, Z, scale, in_cl, first, mo, 0, nullptr, 0.f, 0.f, 0, st);
    const float *sel = meas_in ? meas_in : bound;
    const float ss = meas_in ? 1.f : bscale, th = meas_in ? thr_m : thr;
    int rc = launch_ss_march(vin, vout, B, X, Y, Z, scale, in_cl, first, mo, 0, sel, ss, th, 1, st);
    if (rc) return rc;
    return launch_ss_march(vin, vout, B, X, Y, Z, scale, in_cl, first, mo, 1, sel, ss, th, 2, st);
}

extern "C" int dfm_vecint_fwd(const float *svf, float *out, float *work, int B, int X, int Y, int Z,
                              int nsteps, int save_steps, unsigned flags, void *stream) {
    int rc;
    if ((rc = validate_grid("dfm_vecint_fwd", B, X, Y, Z))) return rc;
    DFM_REQUIRE(nsteps >= 0 && nsteps <= 30, DFM_EINVAL, "dfm_vecint_fwd: nsteps %d out of range [0, 30]", nsteps);
    DFM_REQUIRE(svf && out, DFM_EINVAL, "dfm_vecint_fwd: null pointer");
    DFM_REQUIRE(svf != out, DFM_EINVAL, "dfm_vecint_fwd: out must not alias svf");
    if (B == 0) return DFM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * 3 * X * Y * Z;
    const unsigned in_cl = flags & DFM_FIELD_IN_CL, out_cl = flags & DFM_FIELD_OUT_CL;
    if (nsteps == 0) {   // integrate_vec with nb_steps = 0 is the identity (v / 2**0)
        if (in_cl == (out_cl ? DFM_FIELD_IN_CL : 0u)) {
            cudaError_t e = cudaMemcpyAsync(out, svf, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
            DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_vecint_fwd: %s", cudaGetErrorString(e));
            return DFM_OK;
        }
        return in_cl ? dfm_cl_to_planar(svf, out, B, 3, (size_t)X * Y * Z, 4, stream)
                     : dfm_planar_to_cl(svf, out, B, 3, (size_t)X * Y * Z, 4, stream);
    }
    const size_t need = dfm_vecint_workspace_bytes(B, X, Y, Z, nsteps, save_steps);
    DFM_REQUIRE(need == 0 || work, DFM_EINVAL, "dfm_vecint_fwd: workspace of %zu bytes required", need);
    const float scale0 = ldexpf(1.f, -nsteps);
    // |v_{k+1}| <= 2 max|v_k| (v_{k+1} = v_k + a convex combination of v_k), so one measured maximum
    // bounds every later step: steps whose bound is below the brick's static halo skip the box reduction
    float *bound = work ? work + (save_steps ? (size_t)nsteps : (size_t)1) * n : nullptr;
    // the plane-marching kernel serves every step whose source is one of these three buffers
    const bool march = ss_march_eligible(svf, X, Y, Z) && aligned16(out) && (!work || aligned16(work));
    if (save_steps) {
        // work[k] = v_k (input of step k), k = 0..nsteps-1; v_0 = svf * 2^-nsteps (planar)
        cudaError_t e = cudaMemsetAsync(bound, 0, (size_t)B * sizeof(float), st);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_vecint_fwd: %s", cudaGetErrorString(e));
        rc = scale_copy_to_planar(svf, work, B, (size_t)X * Y * Z, scale0, in_cl != 0, bound, st);   // bound = max|v_0|
        if (rc) return rc;
        for (int k = 0; k < nsteps; ++k) {
            const float *vin = work + (size_t)k * n;
            float *vout = (k == nsteps - 1) ? out : work + (size_t)(k + 1) * n;
            unsigned f = (k == nsteps - 1) ? out_cl : 0u;
            rc = DFM_EUNSUPPORTED;
            if (march && !f) rc = march_step(vin, vout, B, X, Y, Z, 1.f, false, false, nullptr, nsteps - 1 - k, bound, ldexpf(1.f, k), nullptr, nullptr, st);
            if (rc == DFM_EUNSUPPORTED)
                rc = ss_step(vin, vout, B, X, Y, Z, 1.f, f, nsteps - 1 - k, bound, ldexpf(1.f, k), nullptr, st);
            if (rc) return rc;
        }
        return DFM_OK;
    }
    // ping-pong between `work` and `out` so that the last step lands in `out`; intermediates planar.
    // The first step measures max|v_1| per item on the way (marching kernel, or the channels-last kernels);
    // `have_bound` is set only once a kernel that writes it has actually been launched, because the brick
    // kernel's static-halo path trusts the bound without a per-voxel check.
    const bool want_bound = (in_cl || march) && nsteps >= 2;
    bool have_bound = false;
    const size_t bstride = ((size_t)B * sizeof(float) + 255) / 256 * 64;     // floats between the three per-item arrays
    if (want_bound) {
        cudaError_t e = cudaMemsetAsync(bound, 0, 3 * bstride * sizeof(float), st);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_vecint_fwd: %s", cudaGetErrorString(e));
    }
    static const bool no_meas = getenv("DFM_MARCH_NO_MEAS") != nullptr;       // tuning aid: doubled bound only
    const float *cur = svf;
    unsigned cur_cl = in_cl;
    for (int k = 0; k < nsteps; ++k) {
        const bool last = (k == nsteps - 1);
        float *dst = ((nsteps - 1 - k) % 2 == 0) ? out : work;
        unsigned f = cur_cl | (last ? out_cl : 0u);
        const float scale = k == 0 ? scale0 : 1.f;
        float *measure = (want_bound && k == 0) ? bound : nullptr;
        const float *bnd = (have_bound && k >= 1) ? bound : nullptr;       // bnd * 2^(k-1) bounds the input of step k
        rc = DFM_EUNSUPPORTED;
        if (march && !(f & DFM_FIELD_OUT_CL)) {
            // steps nsteps-3 and nsteps-2 record max |out| per item (slots 1 and 2); steps nsteps-2 and nsteps-1 select on them
            const int left = nsteps - 1 - k;
            const bool meas_ok = want_bound && !no_meas && nsteps >= 4;
            float *mout = (meas_ok && k >= 1 && (left == 2 || left == 1)) ? bound + (left == 2 ? 1 : 2) * bstride : nullptr;
            const float *min_ = (meas_ok && k >= 2 && (left == 1 || left == 0)) ? bound + (left == 1 ? 1 : 2) * bstride : nullptr;
            rc = march_step(cur, dst, B, X, Y, Z, scale, cur_cl != 0, k == 0, measure, left, bnd, ldexpf(1.f, k - 1), min_, mout, st);
            if (rc == DFM_OK && measure) have_bound = true;
        }
        if (rc == DFM_EUNSUPPORTED && k == 0 && in_cl && !(f & DFM_FIELD_OUT_CL)) {       // channels-last svf: optimistic static brick
            rc = launch_ss_first_cl(svf, dst, B, X, Y, Z, scale0, measure, st);
            if (rc == DFM_OK && measure) have_bound = true;
        }
        if (rc == DFM_EUNSUPPORTED) {
            rc = ss_step(cur, dst, B, X, Y, Z, scale, f, nsteps - 1 - k, bnd, ldexpf(1.f, k - 1), (in_cl && k == 0) ? measure : nullptr, st);
            if (rc == DFM_OK && in_cl && k == 0 && measure) have_bound = true;   // channels-last input runs the gather kernel, which measures
        }
        if (rc) return rc;
        cur = dst;
        cur_cl = 0u;
    }
    return DFM_OK;
}
