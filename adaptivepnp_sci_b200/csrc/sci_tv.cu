// K4: Chambolle total-variation prior, fused with clip + ADMM dual update.
//
// Replaces the per-iteration  D2H -> skimage.restoration.denoise_tv_chambolle ->
// H2D  round trip of the reference (dvp_linear_inv_2_stage_ADMM_tensor_online.py
// :153-160, :403-407) plus the clip / dual update (:265-267, :501-503).
//
// Algorithm (restated in oracle/tv_chambolle.py): each of the
// 4B channels (frame t, Bayer phase) is an independent (H/2)x(W/2) image that
// lives at stride 2 inside the full-resolution plane t.  n_iter_max = 5 means at
// most 4 dual updates shape the result, so the value at a pixel depends on a
// 4-pixel (half-res) neighbourhood: one block loads a full-res tile plus an
// 8-pixel halo into shared memory and runs ALL inner iterations there (halo
// recompute, no grid sync).  The early stop needs the per-channel energies
// E_0..E_3, which are global sums: the main pass assumes "no early stop" (true
// for ~97% of channel calls), writes deterministic per-block fp64 partial sums,
// a one-block decision kernel derives n_stop per channel, and a fix-up pass
// (same kernel, exits at once unless one of its four channels stopped early)
// rewrites the few channels that did.  No host involvement, 3 launches.
//
// Compiled with --fmad=false: the fp32 arithmetic mirrors numpy's separate ops.
#include <stdlib.h>

#include "sci_common.cuh"

namespace {

constexpr int TV_TW = 64, TV_TH = 32;      // output tile (full-res pixels)
constexpr int TV_HALO = 8;                 // 4 half-res pixels
constexpr int TV_RW = TV_TW + 2 * TV_HALO; // 80
constexpr int TV_RH = TV_TH + 2 * TV_HALO; // 48
constexpr int TV_NPIX = TV_RW * TV_RH;     // 3840
constexpr int TV_ROWGROUPS = 4;            // 80 columns x 4 row groups
constexpr int TV_THREADS = TV_RW * TV_ROWGROUPS;   // 320
constexpr int TV_MAX_UPD = 4;              // out_1..out_4 are the candidate results
#ifndef TV_UNROLL
#define TV_UNROLL 4                        // row-loop unrolling: independent pixels in flight hide the sqrt / divide latency
#endif
constexpr int TV_UNROLL_N = TV_UNROLL;

// workspace: double epart[B][4 iters][2 kinds][4 phases][nblk], then int nstop[B*4]
__host__ __device__ inline size_t tv_epart_count(int B, int nblk) { return (size_t)B * 4 * 2 * 4 * nblk; }

// is_fix = 0: main pass (all channels run to n_iter_max, energies recorded)
// is_fix = 1: fix-up pass (only channels with nstop < last are rewritten)
__global__ void __launch_bounds__(TV_THREADS) tv_chambolle_kernel(
    const float* __restrict__ x, const float* __restrict__ b, float c_b, float* __restrict__ theta,
    float* __restrict__ b_out, float s_b, int clip, int H, int W, float tau, float tw, int last_iter,
    double* __restrict__ epart, const int* __restrict__ nstop, int is_fix) {
    extern __shared__ float smem[];
    float* sf = smem;                 // f = x + c_b*b
    float* so = sf + TV_NPIX;         // current iterate out_i
    float* sp0 = so + TV_NPIX;        // dual variable, row direction
    float* sp1 = sp0 + TV_NPIX;       // dual variable, column direction
    __shared__ double sred[TV_THREADS / 32][16][2];   // 10 warps
    __shared__ int s_stop[4];

    const int t = blockIdx.z;
    const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) s_stop[threadIdx.x] = is_fix ? nstop[t * 4 + threadIdx.x] : last_iter;
    __syncthreads();
    if (is_fix && s_stop[0] >= last_iter && s_stop[1] >= last_iter && s_stop[2] >= last_iter && s_stop[3] >= last_iter)
        return;

    const long plane = (long)H * W;
    const float* xp = x + t * plane;
    const float* bp = b ? b + t * plane : nullptr;
    const int gr0 = blockIdx.y * TV_TH - TV_HALO, gc0 = blockIdx.x * TV_TW - TV_HALO;

    // thread -> fixed column cx of the 80-wide region, rows ry, ry+4, ... (no per-pixel
    // div/mod, column predicates hoisted out of the row loops)
    const int cx = threadIdx.x % TV_RW, ry = threadIdx.x / TV_RW;
    const bool t_active = true;
    const int gc = gc0 + cx;
    const bool col_in = gc >= 0 && gc < W;
    const bool col_interior = cx >= TV_HALO && cx < TV_HALO + TV_TW && gc < W;
    const bool col_has_left = gc >= 2 && cx >= 2;            // neighbour (c-2) exists in the image and in the tile
    const bool col_has_right = gc + 2 < W && cx + 2 < TV_RW;

    if (t_active) {
        for (int rr = ry; rr < TV_RH; rr += TV_ROWGROUPS) {
            const int gr = gr0 + rr, i = rr * TV_RW + cx;
            float v = 0.f;
            if (col_in && gr >= 0 && gr < H) {
                v = xp[(long)gr * W + gc];
                if (bp) v = v + c_b * bp[(long)gr * W + gc];
            }
            sf[i] = v; so[i] = v; sp0[i] = 0.f; sp1[i] = 0.f;
        }
    }
    __syncthreads();

    // energy accumulators: [iteration 0..3][row parity] for (sum d^2, sum |g|); the column parity is fixed per thread
    double e_d[TV_MAX_UPD][2], e_n[TV_MAX_UPD][2];
#pragma unroll
    for (int k = 0; k < TV_MAX_UPD; ++k) { e_d[k][0] = e_d[k][1] = e_n[k][0] = e_n[k][1] = 0.0; }
    const int cpar = gc & 1;

#pragma unroll
    for (int it = 0; it <= TV_MAX_UPD; ++it) {
        if (it > last_iter) break;
        // ---- phase A: d = -div p, out = f + d (it > 0); write the result of the channels stopping here
        if (it > 0) {
            if (t_active) {
#pragma unroll TV_UNROLL_N
                for (int rr = ry; rr < TV_RH; rr += TV_ROWGROUPS) {
                    const int gr = gr0 + rr, i = rr * TV_RW + cx;
                    float d = -(sp0[i] + sp1[i]);
                    if (gr >= 2 && rr >= 2) d += sp0[i - 2 * TV_RW];
                    if (col_has_left) d += sp1[i - 2];
                    const float o = sf[i] + d;
                    so[i] = o;
                    if (col_interior && rr >= TV_HALO && rr < TV_HALO + TV_TH && gr < H) {
                        const int rpar = gr & 1;
                        if (it < TV_MAX_UPD && !is_fix) {
                            const double dd = (double)(d * d);
                            if (rpar) e_d[it][1] += dd; else e_d[it][0] += dd;
                        }
                        const int stop = s_stop[rpar * 2 + cpar];
                        if (it == stop && (!is_fix || stop < last_iter)) {
                            const long g = (long)gr * W + gc;
                            float th = o;
                            if (clip) th = fminf(fmaxf(th, 0.f), 1.f);
                            theta[t * plane + g] = th;
                            if (bp) b_out[t * plane + g] = bp[g] + s_b * (xp[g] - th);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (it == TV_MAX_UPD || it == last_iter) break;   // the last dual update never shapes the result
        // ---- phase B: forward differences, energy, dual update
        if (t_active) {
#pragma unroll TV_UNROLL_N
            for (int rr = ry; rr < TV_RH; rr += TV_ROWGROUPS) {
                const int gr = gr0 + rr, i = rr * TV_RW + cx;
                const bool in_img = col_in && gr >= 0 && gr < H;
                const float o = so[i];
                float g0 = 0.f, g1 = 0.f;
                if (in_img && gr + 2 < H && rr + 2 < TV_RH) g0 = so[i + 2 * TV_RW] - o;
                if (in_img && col_has_right) g1 = so[i + 2] - o;
                const float nrm = sqrtf(g0 * g0 + g1 * g1);
                if (col_interior && rr >= TV_HALO && rr < TV_HALO + TV_TH && gr < H && !is_fix) {
                    if (gr & 1) e_n[it][1] += (double)nrm; else e_n[it][0] += (double)nrm;
                }
                const float inv = 1.0f / (nrm * tw + 1.0f);          // one IEEE division shared by both components
                sp0[i] = in_img ? (sp0[i] - tau * g0) * inv : 0.f;
                sp1[i] = in_img ? (sp1[i] - tau * g1) * inv : 0.f;
            }
        }
        __syncthreads();
    }

    if (is_fix) return;
    // ---- deterministic per-block energy partials: warp shuffle (keeping lane parity = column parity), then
    //      a fixed-order sum over the 8 warps
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < TV_MAX_UPD; ++k)
#pragma unroll
        for (int rp = 0; rp < 2; ++rp) {
            double a = e_d[k][rp], c = e_n[k][rp];
#pragma unroll
            for (int o = 16; o >= 2; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if (lane < 2) {      // lane parity == thread parity == column parity offset
                sred[wid][(k * 2 + rp) * 2 + 0][lane] = a;
                sred[wid][(k * 2 + rp) * 2 + 1][lane] = c;
            }
        }
    __syncthreads();
    if (threadIdx.x < 32) {
        // thread -> (k, rp, kind, lane_par)
        const int lp = threadIdx.x & 1, kind = (threadIdx.x >> 1) & 1, rp = (threadIdx.x >> 2) & 1, k = threadIdx.x >> 3;
        double s = 0.0;
        for (int w8 = 0; w8 < TV_THREADS / 32; ++w8) s += sred[w8][(k * 2 + rp) * 2 + kind][lp];
        // lane parity lp == column parity: cx = threadIdx.x % 80 keeps the parity of threadIdx.x, gc0 is even
        const int cp = (gc0 + lp) & 1;
        const int phase = rp * 2 + cp;
        epart[((((size_t)t * 4 + k) * 2 + kind) * 4 + phase) * nblk + blk] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// Version 2 of the tile kernel (same maths, same interface, bit-identical interior results).  ncu on version 1
// (profiles/ncu_r1_prof_tv_chambolle.csv): issue slots 65 % busy, DRAM 2 % — instruction-bound, and most instructions were
// boundary predicates and address arithmetic.  Here
//   * the four shared-memory arrays carry a 2-pixel zero border, so tile-edge neighbours need no predicate (halo pixels
//     may then hold garbage one ring deeper per iteration, which the 8-pixel halo already budgets for);
//   * image borders are handled by per-row / per-thread select masks (dual variables outside the image stay exactly 0);
//   * a thread owns a PAIR of adjacent columns (the two column phases) of rows ry, ry+8, ..: every shared-memory access is
//     one 64-bit load/store for two pixels and the address arithmetic is shared.
// ---------------------------------------------------------------------------------------------------
constexpr int TV2_P = TV_RW + 4;                   // padded row pitch (84, even: float2 aligned)
constexpr int TV2_ROWS = TV_RH + 4;                // 52
constexpr int TV2_N = TV2_P * TV2_ROWS;            // 4368 floats per array
constexpr int TV2_PAIRS = TV_RW / 2;               // 40 column pairs
constexpr int TV2_RG = TV_THREADS / TV2_PAIRS;     // 8 row groups
constexpr int TV2_RPT = TV_RH / TV2_RG;            // 6 rows per thread

__global__ void __launch_bounds__(TV_THREADS, 3) tv_chambolle2_kernel(
    const float* __restrict__ x, const float* __restrict__ b, float c_b, float* __restrict__ theta,
    float* __restrict__ b_out, float s_b, int clip, int H, int W, float tau, float tw, int last_iter,
    double* __restrict__ epart, const int* __restrict__ nstop, int is_fix) {
    extern __shared__ float smem[];
    float* sf = smem;
    float* so = sf + TV2_N;
    float* sp0 = so + TV2_N;
    float* sp1 = sp0 + TV2_N;
    __shared__ int s_stop[4];

    const int t = blockIdx.z;
    const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) s_stop[threadIdx.x] = is_fix ? nstop[t * 4 + threadIdx.x] : last_iter;
    __syncthreads();
    if (is_fix && s_stop[0] >= last_iter && s_stop[1] >= last_iter && s_stop[2] >= last_iter && s_stop[3] >= last_iter)
        return;

    const long plane = (long)H * W;
    const float* xp = x + t * plane;
    const float* bp = b ? b + t * plane : nullptr;
    const int gr0 = blockIdx.y * TV_TH - TV_HALO, gc0 = blockIdx.x * TV_TW - TV_HALO;    // both even

    // zero everything once (borders stay zero for the whole kernel)
    {
        float4* z = reinterpret_cast<float4*>(smem);
        for (int i = threadIdx.x; i < 4 * TV2_N / 4; i += TV_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const int cx = (threadIdx.x % TV2_PAIRS) * 2, ry = threadIdx.x / TV2_PAIRS;
    const int gc = gc0 + cx;                                  // even; the pair is (gc, gc+1) = column phases 0, 1
    const bool col_in = gc >= 0 && gc < W;                    // W even: both columns in or both out
    const bool col_right = col_in && gc + 2 < W;              // forward neighbour (c+2) exists in the image, for both
    const bool col_interior = cx >= TV_HALO && cx < TV_HALO + TV_TW && gc < W;
    const int rpar = ry & 1;                                  // gr0 even, rows advance by 8: fixed row parity per thread
    const int base = (ry + 2) * TV2_P + cx + 2;               // smem index of (rr = ry, cx)

    // ---- load f = x + c_b*b
#pragma unroll
    for (int k = 0; k < TV2_RPT; ++k) {
        const int gr = gr0 + ry + k * TV2_RG, i = base + k * TV2_RG * TV2_P;
        float2 v = make_float2(0.f, 0.f);
        if (col_in && gr >= 0 && gr < H) {
            v = *reinterpret_cast<const float2*>(xp + (long)gr * W + gc);
            if (bp) {
                const float2 bb = *reinterpret_cast<const float2*>(bp + (long)gr * W + gc);
                v.x = v.x + c_b * bb.x; v.y = v.y + c_b * bb.y;
            }
        }
        *reinterpret_cast<float2*>(sf + i) = v;
        *reinterpret_cast<float2*>(so + i) = v;
    }
    __syncthreads();

    double e_d[TV_MAX_UPD][2], e_n[TV_MAX_UPD][2];            // [iteration][column phase]
#pragma unroll
    for (int k = 0; k < TV_MAX_UPD; ++k) { e_d[k][0] = e_d[k][1] = e_n[k][0] = e_n[k][1] = 0.0; }
    const int stop0 = s_stop[rpar * 2 + 0], stop1 = s_stop[rpar * 2 + 1];

#pragma unroll
    for (int it = 0; it <= TV_MAX_UPD; ++it) {
        if (it > last_iter) break;
        if (it > 0) {
            // ---- phase A: d = -div p, out = f + d; write the result of the channels stopping here
#pragma unroll 2
            for (int k = 0; k < TV2_RPT; ++k) {
                const int rr = ry + k * TV2_RG, gr = gr0 + rr, i = base + k * TV2_RG * TV2_P;
                const float2 p0 = *reinterpret_cast<const float2*>(sp0 + i), p1 = *reinterpret_cast<const float2*>(sp1 + i);
                const float2 pu = *reinterpret_cast<const float2*>(sp0 + i - 2 * TV2_P);
                const float2 pl = *reinterpret_cast<const float2*>(sp1 + i - 2);
                const float2 f = *reinterpret_cast<const float2*>(sf + i);
                float dx = -(p0.x + p1.x), dy = -(p0.y + p1.y);
                dx += pu.x; dy += pu.y;
                dx += pl.x; dy += pl.y;
                const float ox = f.x + dx, oy = f.y + dy;
                *reinterpret_cast<float2*>(so + i) = make_float2(ox, oy);
                if (col_interior && rr >= TV_HALO && rr < TV_HALO + TV_TH && gr < H) {
                    if (it < TV_MAX_UPD && !is_fix) {
                        e_d[it][0] += (double)(dx * dx);
                        e_d[it][1] += (double)(dy * dy);
                    }
                    const long g = (long)gr * W + gc;
                    const bool w0 = it == stop0 && (!is_fix || stop0 < last_iter);
                    const bool w1 = it == stop1 && (!is_fix || stop1 < last_iter);
                    if (w0 || w1) {
                        float t0 = ox, t1 = oy;
                        if (clip) { t0 = fminf(fmaxf(t0, 0.f), 1.f); t1 = fminf(fmaxf(t1, 0.f), 1.f); }
                        if (w0 && w1) {
                            *reinterpret_cast<float2*>(theta + t * plane + g) = make_float2(t0, t1);
                            if (bp) {
                                const float2 bb = *reinterpret_cast<const float2*>(bp + g), xx = *reinterpret_cast<const float2*>(xp + g);
                                *reinterpret_cast<float2*>(b_out + t * plane + g) =
                                    make_float2(bb.x + s_b * (xx.x - t0), bb.y + s_b * (xx.y - t1));
                            }
                        } else {
                            const int j = w0 ? 0 : 1;
                            const float th = w0 ? t0 : t1;
                            theta[t * plane + g + j] = th;
                            if (bp) b_out[t * plane + g + j] = bp[g + j] + s_b * (xp[g + j] - th);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (it == TV_MAX_UPD || it == last_iter) break;       // the last dual update never shapes the result
        // ---- phase B: forward differences, energy, dual update
#pragma unroll 2
        for (int k = 0; k < TV2_RPT; ++k) {
            const int rr = ry + k * TV2_RG, gr = gr0 + rr, i = base + k * TV2_RG * TV2_P;
            const bool in_img = col_in && gr >= 0 && gr < H;
            const bool has_dn = in_img && gr + 2 < H, has_rt = in_img && col_right;
            const float2 o = *reinterpret_cast<const float2*>(so + i);
            const float2 od = *reinterpret_cast<const float2*>(so + i + 2 * TV2_P);
            const float2 orr = *reinterpret_cast<const float2*>(so + i + 2);
            const float g0x = has_dn ? od.x - o.x : 0.f, g0y = has_dn ? od.y - o.y : 0.f;
            const float g1x = has_rt ? orr.x - o.x : 0.f, g1y = has_rt ? orr.y - o.y : 0.f;
            const float nx = sqrtf(g0x * g0x + g1x * g1x), ny = sqrtf(g0y * g0y + g1y * g1y);
            if (col_interior && rr >= TV_HALO && rr < TV_HALO + TV_TH && gr < H && !is_fix) {
                e_n[it][0] += (double)nx;
                e_n[it][1] += (double)ny;
            }
            const float ix = 1.0f / (nx * tw + 1.0f), iy = 1.0f / (ny * tw + 1.0f);
            const float2 p0 = *reinterpret_cast<const float2*>(sp0 + i), p1 = *reinterpret_cast<const float2*>(sp1 + i);
            *reinterpret_cast<float2*>(sp0 + i) = in_img ? make_float2((p0.x - tau * g0x) * ix, (p0.y - tau * g0y) * iy)
                                                         : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(sp1 + i) = in_img ? make_float2((p1.x - tau * g1x) * ix, (p1.y - tau * g1y) * iy)
                                                         : make_float2(0.f, 0.f);
        }
        __syncthreads();
    }

    if (is_fix) return;
    // ---- deterministic per-block energy partials: every thread parks its 16 sums in shared memory (the image arrays
    //      are dead now), 32 threads add them up in a fixed order
    __syncthreads();
    double* sd = reinterpret_cast<double*>(smem);             // [TV_THREADS][16]
#pragma unroll
    for (int k = 0; k < TV_MAX_UPD; ++k) {
        sd[threadIdx.x * 16 + k * 4 + 0] = e_d[k][0];
        sd[threadIdx.x * 16 + k * 4 + 1] = e_d[k][1];
        sd[threadIdx.x * 16 + k * 4 + 2] = e_n[k][0];
        sd[threadIdx.x * 16 + k * 4 + 3] = e_n[k][1];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // thread -> (k, kind, row parity rp, column phase cp)
        const int cp = threadIdx.x & 1, rp = (threadIdx.x >> 1) & 1, kind = (threadIdx.x >> 2) & 1, k = threadIdx.x >> 3;
        double s = 0.0;
        for (int g = rp; g < TV2_RG; g += 2)                  // row groups with this row parity
            for (int c = 0; c < TV2_PAIRS; ++c) s += sd[(g * TV2_PAIRS + c) * 16 + k * 4 + kind * 2 + cp];
        epart[((((size_t)t * 4 + k) * 2 + kind) * 4 + (rp * 2 + cp)) * nblk + blk] = s;
    }
}

// One block per channel: fixed-order (deterministic) reduction of the per-block partials, then the
// reference's stopping rule  |E_prev - E_i| < eps * E_init  (i >= 1).
constexpr int TV_DEC_THREADS = 128;
__global__ void __launch_bounds__(TV_DEC_THREADS) tv_decide_kernel(const double* __restrict__ epart, int nblk, double weight,
                                                                   double eps, double n_pix, int last_iter,
                                                                   int* __restrict__ nstop, int* __restrict__ nstop_out) {
    __shared__ double sm[TV_DEC_THREADS][8];
    const int ch = blockIdx.x, t = ch >> 2, phase = ch & 3;
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {        // q = k*2 + kind
        const double* p = epart + ((((size_t)t * 4 + (q >> 1)) * 2 + (q & 1)) * 4 + phase) * nblk;
        double s = 0.0;
        for (int j = threadIdx.x; j < nblk; j += TV_DEC_THREADS) s += p[j];
        acc[q] = s;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) sm[threadIdx.x][q] = acc[q];
    __syncthreads();
    for (int off = TV_DEC_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + off][q];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double E_init = 0.0, E_prev = 0.0;
        int stop = last_iter;
        for (int k = 0; k < TV_MAX_UPD && k < last_iter; ++k) {
            double E = sm[0][k * 2 + 0];
            E += weight * sm[0][k * 2 + 1];
            E /= n_pix;
            if (k == 0) { E_init = E; E_prev = E; }
            else if (fabs(E_prev - E) < eps * E_init) { stop = k; break; }
            else E_prev = E;
        }
        nstop[ch] = stop;
        if (nstop_out) nstop_out[ch] = stop;
    }
}

}  // namespace

extern "C" size_t sci_tv_workspace_bytes(int H, int W, int B) {
    if (H <= 0 || W <= 0 || B <= 0) return 0;
    const int nblk = sci_ceil_div(W, TV_TW) * sci_ceil_div(H, TV_TH);
    return tv_epart_count(B, nblk) * sizeof(double) + (size_t)B * 4 * sizeof(int);
}

extern "C" int sci_tv_chambolle2d(const float* x, const float* b, float c_b, float* theta, float* b_out, float s_b,
                                  int clip, int H, int W, int B, float weight, float eps, int n_iter_max,
                                  void* workspace, size_t workspace_bytes, int* nstop_out, void* stream) {
    SCI_REQUIRE(x && theta && workspace, "tv: null pointer");
    SCI_REQUIRE((b == nullptr) == (b_out == nullptr), "tv: b and b_out go together");
    SCI_REQUIRE(b == nullptr || b != b_out, "tv: b_out must not alias b");
    SCI_REQUIRE(H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && B > 0 && B <= 65535, "tv: shape");
    if (n_iter_max < 1 || n_iter_max > TV_MAX_UPD + 1)
        return sci_fail(SCI_EUNSUPPORTED, "tv: n_iter_max must be in 1..5 (the reference uses 5)");
    if (workspace_bytes < sci_tv_workspace_bytes(H, W, B)) return sci_fail(SCI_EWORKSPACE, "tv: workspace too small");
    cudaStream_t st = sci_stream(stream);
    const dim3 grid(sci_ceil_div(W, TV_TW), sci_ceil_div(H, TV_TH), B);
    const int nblk = grid.x * grid.y;
    double* epart = reinterpret_cast<double*>(workspace);
    int* nstop = reinterpret_cast<int*>(epart + tv_epart_count(B, nblk));
    const char* v2env = getenv("SCI_TV_V2");
    const bool v2 = !(v2env && v2env[0] == '0');
    auto kern = v2 ? tv_chambolle2_kernel : tv_chambolle_kernel;
    const size_t smem = v2 ? (size_t)4 * TV2_N * sizeof(float) : (size_t)4 * TV_NPIX * sizeof(float);
    static bool attr_set[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev][v2]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return sci_fail(SCI_ELAUNCH, "tv: smem attribute", e);
        if (dev >= 0 && dev < 64) attr_set[dev][v2] = true;
    }
    const int last_iter = n_iter_max - 1;              // index of the last iterate (4 for n_iter_max=5)
    const float tau = 0.25f;                            // 1/(2*ndim), ndim = 2
    const float tw = (float)(0.25 / (double)weight);    // python-double tau/weight, cast at the fp32 multiply
    if (last_iter == 0) {
        // n_iter_max = 1 returns the input unchanged
        return sci_fail(SCI_EUNSUPPORTED, "tv: n_iter_max = 1 is the identity; not routed through the kernel");
    }
    kern<<<grid, TV_THREADS, smem, st>>>(x, b, c_b, theta, b_out, s_b, clip, H, W, tau, tw, last_iter, epart, nullptr, 0);
    SCI_CHECK_LAUNCH("tv main pass");
    const int nch = B * 4;
    tv_decide_kernel<<<nch, TV_DEC_THREADS, 0, st>>>(epart, nblk, (double)weight, (double)eps,
                                                     (double)(H / 2) * (double)(W / 2), last_iter, nstop, nstop_out);
    SCI_CHECK_LAUNCH("tv decide");
    kern<<<grid, TV_THREADS, smem, st>>>(x, b, c_b, theta, b_out, s_b, clip, H, W, tau, tw, last_iter, epart, nstop, 1);
    SCI_CHECK_LAUNCH("tv fix-up pass");
    return SCI_OK;
}
