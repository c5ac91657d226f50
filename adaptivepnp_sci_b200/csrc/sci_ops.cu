// HBM-bound kernels of the ADMM loop: layout remaps, Bayer split/init, sensing
// operators, fused Euclidean projection, Malvar-2004 demosaic, Bayer sampling +
// dual updates, PSNR accumulation.  Compiled with --fmad=false so that the fp32
// arithmetic follows the reference's separate multiply/add ATen ops.
#include "sci_common.cuh"

thread_local char g_sci_last_error[256] = "";

extern "C" int sci_version(void) { return 1000; }
extern "C" const char* sci_last_error(void) { return g_sci_last_error; }

// ---------------------------------------------------------------------------
// Layout remaps: in[P][K] (K = C*B, index c*B+t)  <->  out[B][C][P].
// One block moves 64 pixels through a padded shared tile so that both the
// global read and the global write are fully coalesced.
// ---------------------------------------------------------------------------
constexpr int REMAP_PIX = 64;

__global__ void __launch_bounds__(256) pixlast_to_planar_kernel(const float* __restrict__ in,
                                                                 float* __restrict__ out, int P, int C, int B) {
    extern __shared__ float tile[];
    const int K = C * B, KS = K | 1;
    const long p0 = (long)blockIdx.x * REMAP_PIX;
    const int np = min((long)REMAP_PIX, P - p0);
    const float* src = in + p0 * K;
    for (int i = threadIdx.x; i < np * K; i += blockDim.x) tile[(i / K) * KS + (i % K)] = src[i];
    __syncthreads();
    for (int i = threadIdx.x; i < K * REMAP_PIX; i += blockDim.x) {
        const int k = i / REMAP_PIX, p = i % REMAP_PIX;       // k = c*B + t
        if (p < np) {
            const int c = k / B, t = k % B;
            out[((long)t * C + c) * P + p0 + p] = tile[p * KS + k];
        }
    }
}

__global__ void __launch_bounds__(256) planar_to_pixlast_kernel(const float* __restrict__ in,
                                                                 float* __restrict__ out, int P, int C, int B) {
    extern __shared__ float tile[];
    const int K = C * B, KS = K | 1;
    const long p0 = (long)blockIdx.x * REMAP_PIX;
    const int np = min((long)REMAP_PIX, P - p0);
    for (int i = threadIdx.x; i < K * REMAP_PIX; i += blockDim.x) {
        const int k = i / REMAP_PIX, p = i % REMAP_PIX;
        if (p < np) {
            const int c = k / B, t = k % B;
            tile[p * KS + k] = in[((long)t * C + c) * P + p0 + p];
        }
    }
    __syncthreads();
    float* dst = out + p0 * K;
    for (int i = threadIdx.x; i < np * K; i += blockDim.x) dst[i] = tile[(i / K) * KS + (i % K)];
}

static int remap_launch(bool to_planar, const float* in, float* out, int P, int C, int B, void* stream) {
    SCI_REQUIRE(in && out && P > 0 && C > 0 && B > 0, "remap");
    const int K = C * B;
    if (K > 384) return sci_fail(SCI_EUNSUPPORTED, "remap: C*B > 384");
    const size_t smem = (size_t)REMAP_PIX * (K | 1) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(pixlast_to_planar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(planar_to_pixlast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int grid = sci_ceil_div(P, REMAP_PIX);
    if (to_planar) pixlast_to_planar_kernel<<<grid, 256, smem, sci_stream(stream)>>>(in, out, P, C, B);
    else           planar_to_pixlast_kernel<<<grid, 256, smem, sci_stream(stream)>>>(in, out, P, C, B);
    SCI_CHECK_LAUNCH("remap");
    return SCI_OK;
}

extern "C" int sci_pixlast_to_planar(const float* in, float* out, int P, int C, int B, void* stream) {
    return remap_launch(true, in, out, P, C, B, stream);
}
extern "C" int sci_planar_to_pixlast(const float* in, float* out, int P, int C, int B, void* stream) {
    return remap_launch(false, in, out, P, C, B, stream);
}

// ---------------------------------------------------------------------------
// K0: Bayer split + Phi-sum + init.  Same tile transpose as above, with the
// per-pixel mask sum (zeros -> 1) and the At(y,Phi) / warm-start initialisation
// done while the tile is in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_init_kernel(const float* __restrict__ y, const float* __restrict__ phi_hwb,
                                                          const float* __restrict__ x0_hwb, float* __restrict__ phi,
                                                          float* __restrict__ phisum, float* __restrict__ theta0,
                                                          int P, int B) {
    extern __shared__ float tile[];
    const int KS = B | 1;
    float* tphi = tile;
    float* tx0 = tile + REMAP_PIX * KS;
    const long p0 = (long)blockIdx.x * REMAP_PIX;
    const int np = min((long)REMAP_PIX, P - p0);
    for (int i = threadIdx.x; i < np * B; i += blockDim.x) {
        tphi[(i / B) * KS + (i % B)] = phi_hwb[p0 * B + i];
        if (x0_hwb) tx0[(i / B) * KS + (i % B)] = x0_hwb[p0 * B + i];
    }
    __syncthreads();
    if (threadIdx.x < np) {
        float s = 0.f;
        for (int t = 0; t < B; ++t) s += tphi[threadIdx.x * KS + t];
        phisum[p0 + threadIdx.x] = (s == 0.f) ? 1.f : s;
    }
    for (int i = threadIdx.x; i < B * REMAP_PIX; i += blockDim.x) {
        const int t = i / REMAP_PIX, p = i % REMAP_PIX;
        if (p < np) {
            const float ph = tphi[p * KS + t];
            phi[(long)t * P + p0 + p] = ph;
            theta0[(long)t * P + p0 + p] = x0_hwb ? tx0[p * KS + t] : y[p0 + p] * ph;
        }
    }
}

extern "C" int sci_bayer_split_init(const float* y, const float* phi_hwb, const float* x0_hwb, float* phi,
                                    float* phisum, float* theta0, int H, int W, int B, void* stream) {
    SCI_REQUIRE(y && phi_hwb && phi && phisum && theta0, "split_init: null pointer");
    SCI_REQUIRE(H > 0 && W > 0 && (H % 2 == 0) && (W % 2 == 0) && B > 0 && B <= 512, "split_init: shape");
    const int P = H * W;
    const size_t smem = (size_t)2 * REMAP_PIX * (B | 1) * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(split_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    split_init_kernel<<<sci_ceil_div(P, REMAP_PIX), 256, smem, sci_stream(stream)>>>(y, phi_hwb, x0_hwb, phi, phisum,
                                                                                     theta0, P, B);
    SCI_CHECK_LAUNCH("split_init");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// A_/At_ on arbitrary-stride views (API parity with utilspy.py:28-44).
// ---------------------------------------------------------------------------
__global__ void A_strided_kernel(const float* __restrict__ x, long sxh, long sxw, long sxt,
                                 const float* __restrict__ phi, long sph, long spw, long spt,
                                 float* __restrict__ y, long syh, long syw, int h, int w, int B) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= w) return;
    const float* xp = x + i * sxh + j * sxw;
    const float* pp = phi + i * sph + j * spw;
    float acc = 0.f;
    for (int t = 0; t < B; ++t) acc += xp[t * sxt] * pp[t * spt];
    y[i * syh + j * syw] = acc;
}

__global__ void At_strided_kernel(const float* __restrict__ y, long syh, long syw,
                                  const float* __restrict__ phi, long sph, long spw, long spt,
                                  float* __restrict__ x, long sxh, long sxw, long sxt, int h, int w, int B) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= w) return;
    const float yv = y[i * syh + j * syw];
    const float* pp = phi + i * sph + j * spw;
    float* xp = x + i * sxh + j * sxw;
    for (int t = 0; t < B; ++t) xp[t * sxt] = yv * pp[t * spt];
}

extern "C" int sci_A(const float* x, long sxh, long sxw, long sxt, const float* phi, long sph, long spw, long spt,
                     float* y, long syh, long syw, int h, int w, int B, void* stream) {
    SCI_REQUIRE(x && phi && y && h > 0 && w > 0 && B > 0 && h <= 65535, "sci_A");
    A_strided_kernel<<<dim3(sci_ceil_div(w, 128), h), 128, 0, sci_stream(stream)>>>(x, sxh, sxw, sxt, phi, sph, spw, spt,
                                                                                     y, syh, syw, h, w, B);
    SCI_CHECK_LAUNCH("sci_A");
    return SCI_OK;
}

extern "C" int sci_At(const float* y, long syh, long syw, const float* phi, long sph, long spw, long spt, float* x,
                      long sxh, long sxw, long sxt, int h, int w, int B, void* stream) {
    SCI_REQUIRE(x && phi && y && h > 0 && w > 0 && B > 0 && h <= 65535, "sci_At");
    At_strided_kernel<<<dim3(sci_ceil_div(w, 128), h), 128, 0, sci_stream(stream)>>>(y, syh, syw, phi, sph, spw, spt, x,
                                                                                      sxh, sxw, sxt, h, w, B);
    SCI_CHECK_LAUNCH("sci_At");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// K1/K2: fused projection.  v = theta + c_b*b ; r = (y - sum_t phi*v)/(phisum + c_den) ;
// x = v + c_l*(phi*r).   One thread owns 4 consecutive pixels (float4) of every
// frame: 3*B independent 128-bit loads in flight, v and phi stay in registers,
// a single pass over the cubes (algorithmic 4 cubes + 2 planes).
// ---------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&r)[VEC]) {
    *reinterpret_cast<typename VecT<VEC>::type*>(r) = *reinterpret_cast<const typename VecT<VEC>::type*>(p);
}
template <int VEC>
__device__ __forceinline__ void load_vec_stream(const float* p, float (&r)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = ldg_stream4(p);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else {
        *reinterpret_cast<typename VecT<VEC>::type*>(r) = __ldg(reinterpret_cast<const typename VecT<VEC>::type*>(p));
    }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&r)[VEC]) {
    *reinterpret_cast<typename VecT<VEC>::type*>(p) = *reinterpret_cast<const typename VecT<VEC>::type*>(r);
}

// One thread owns VEC consecutive pixels of every frame: 3*B independent vector loads in flight,
// v and phi stay in registers, a single pass over the cubes.  VEC is chosen by the launcher so that
// small cubes still fill the 148 SMs (more, narrower threads) and large ones use 128-bit accesses.
template <int B, int VEC>
__global__ void __launch_bounds__(256) project_kernel_vec(const float* __restrict__ theta, const float* __restrict__ b,
                                                           const float* __restrict__ phi, const float* __restrict__ y,
                                                           const float* __restrict__ phisum, float* __restrict__ x,
                                                           long npix, float c_b, float c_l, float c_den,
                                                           const float* __restrict__ orig, double* __restrict__ sse) {
    __shared__ double red[32];
    const long q = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    double err = 0.0;
    if (q < npix) {
        float v[B][VEC], ph[B][VEC];
#pragma unroll
        for (int t = 0; t < B; ++t) {
            float th[VEC], bb[VEC];
            load_vec<VEC>(theta + t * npix + q, th);
            load_vec<VEC>(b + t * npix + q, bb);
            load_vec_stream<VEC>(phi + t * npix + q, ph[t]);
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[t][e] = th[e] + c_b * bb[e];
        }
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int t = 0; t < B; ++t)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += v[t][e] * ph[t][e];
        float yy[VEC], ps[VEC], r[VEC];
        load_vec_stream<VEC>(y + q, yy);
        load_vec_stream<VEC>(phisum + q, ps);
#pragma unroll
        for (int e = 0; e < VEC; ++e) r[e] = (yy[e] - acc[e]) / (ps[e] + c_den);
#pragma unroll
        for (int t = 0; t < B; ++t) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[e] = v[t][e] + c_l * (ph[t][e] * r[e]);
            store_vec<VEC>(x + t * npix + q, o);
            if (orig) {
                float og[VEC];
                load_vec_stream<VEC>(orig + t * npix + q, og);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { const float d = o[e] - og[e]; err += (double)(d * d); }
            }
        }
    }
    if (orig) {
        const double s = block_sum(err, red);
        if (threadIdx.x == 0) atomicAdd(sse, s);
    }
}

// Generic frame count / unaligned pixel count: one pixel per thread, two passes
// over the frames (second pass re-reads theta, b, phi through L1/L2).
__global__ void __launch_bounds__(256) project_kernel_generic(const float* __restrict__ theta, const float* __restrict__ b,
                                                               const float* __restrict__ phi, const float* __restrict__ y,
                                                               const float* __restrict__ phisum, float* __restrict__ x,
                                                               long npix, int B, float c_b, float c_l, float c_den,
                                                               const float* __restrict__ orig, double* __restrict__ sse) {
    __shared__ double red[32];
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (q < npix) {
        float acc = 0.f;
        for (int t = 0; t < B; ++t) acc += (theta[t * npix + q] + c_b * b[t * npix + q]) * phi[t * npix + q];
        const float r = (y[q] - acc) / (phisum[q] + c_den);
        for (int t = 0; t < B; ++t) {
            const float v = theta[t * npix + q] + c_b * b[t * npix + q];
            const float o = v + c_l * (phi[t * npix + q] * r);
            x[t * npix + q] = o;
            if (orig) { const float d = o - orig[t * npix + q]; err += (double)(d * d); }
        }
    }
    if (orig) {
        const double s = block_sum(err, red);
        if (threadIdx.x == 0) atomicAdd(sse, s);
    }
}

static int project_launch(const float* theta, const float* b, const float* phi, const float* y, const float* phisum,
                          float* x, long npix, int B, float c_b, float c_l, float c_den, const float* orig, double* sse,
                          void* stream) {
    SCI_REQUIRE(theta && b && phi && y && phisum && x && npix > 0 && B > 0, "project: null/shape");
    SCI_REQUIRE(!orig || sse, "project: orig given without sse");
    cudaStream_t st = sci_stream(stream);
    const bool aligned = (npix % 4 == 0) && ((((uintptr_t)theta | (uintptr_t)b | (uintptr_t)phi | (uintptr_t)y |
                                               (uintptr_t)phisum | (uintptr_t)x | (uintptr_t)orig) & 15) == 0);
    // narrower threads for small cubes: keep >= ~4 blocks per SM in flight
    const long want_threads = (long)SCI_NUM_SMS * 4 * 256;
    const int vecw = (npix / 4 >= want_threads) ? 4 : (npix / 2 >= want_threads ? 2 : 1);
#define SCI_PROJ_LAUNCH(BB, VV) project_kernel_vec<BB, VV><<<sci_ceil_div(npix / VV, 256), 256, 0, st>>>( \
        theta, b, phi, y, phisum, x, npix, c_b, c_l, c_den, orig, sse)
#define SCI_PROJ_CASE(BB) case BB: if (vecw == 4) SCI_PROJ_LAUNCH(BB, 4); else if (vecw == 2) SCI_PROJ_LAUNCH(BB, 2); \
                                   else SCI_PROJ_LAUNCH(BB, 1); break;
    bool done = false;
    if (aligned) {
        done = true;
        switch (B) {
            SCI_PROJ_CASE(4) SCI_PROJ_CASE(8) SCI_PROJ_CASE(10) SCI_PROJ_CASE(12) SCI_PROJ_CASE(16)
            case 24: if (vecw >= 2) SCI_PROJ_LAUNCH(24, 2); else SCI_PROJ_LAUNCH(24, 1); break;
            case 32: SCI_PROJ_LAUNCH(32, 1); break;
            default: done = false;
        }
    }
#undef SCI_PROJ_CASE
#undef SCI_PROJ_LAUNCH
    if (!done)
        project_kernel_generic<<<sci_ceil_div(npix, 256), 256, 0, st>>>(theta, b, phi, y, phisum, x, npix, B, c_b, c_l,
                                                                         c_den, orig, sse);
    SCI_CHECK_LAUNCH("project");
    return SCI_OK;
}

extern "C" int sci_project_stage1(const float* theta, const float* b, const float* phi, const float* y,
                                  const float* phisum, float* x, long npix, int B, float lambda_, float gamma,
                                  const float* orig, double* sse, void* stream) {
    return project_launch(theta, b, phi, y, phisum, x, npix, B, 1.0f, lambda_, gamma, orig, sse, stream);
}

extern "C" int sci_project_stage2(const float* theta, const float* b, const float* phi, const float* y,
                                  const float* phisum, float* x, long npix, int B, double alpha, double rho, void* stream) {
    // (1/rou) and alpha*rou are python doubles in the reference, cast to fp32 at the tensor op: rho arrives as a double so
    // that float32(1/0.55) = 1.8181819f is formed exactly as there (a float rho gave 1.8181818f, 1 ulp off)
    const float c_b = -(float)(1.0 / rho);
    const float c_den = (float)(alpha * rho);
    return project_launch(theta, b, phi, y, phisum, x, npix, B, c_b, 1.0f, c_den, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------
// K6: Malvar-2004 on the mosaic m = x + c_b*b, all frames in one launch.
// 32x8 output tile + 2-pixel halo in shared memory, torch 'reflect' borders.
// Each thread evaluates only the two 5x5 correlations its CFA site needs.
// ---------------------------------------------------------------------------
constexpr int MV_TW = 32, MV_TH = 8, MV_SW = MV_TW + 4, MV_SH = MV_TH + 4;

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void __launch_bounds__(MV_TW * MV_TH) malvar_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                               float c_b, const float* __restrict__ w, float inv_tau,
                                                               float* __restrict__ x_rgb, float* __restrict__ u, int H,
                                                               int W) {
    __shared__ float s[MV_SH][MV_SW + 1];
    const int t = blockIdx.z;
    const long plane = (long)H * W;
    const float* xp = x + t * plane;
    const float* bp = b ? b + t * plane : nullptr;
    const int c0 = blockIdx.x * MV_TW - 2, r0 = blockIdx.y * MV_TH - 2;
    for (int i = threadIdx.y * MV_TW + threadIdx.x; i < MV_SH * MV_SW; i += MV_TW * MV_TH) {
        const int rr = i / MV_SW, cc = i % MV_SW;
        const int gr = reflect_idx(min(r0 + rr, H + 1), H), gc = reflect_idx(min(c0 + cc, W + 1), W);
        float v = xp[(long)gr * W + gc];
        if (bp) v = v + c_b * bp[(long)gr * W + gc];
        s[rr][cc] = v;
    }
    __syncthreads();
    const int col = blockIdx.x * MV_TW + threadIdx.x, row = blockIdx.y * MV_TH + threadIdx.y;
    if (col >= W || row >= H) return;
    const int y0 = threadIdx.y + 2, x0 = threadIdx.x + 2;
#define S(dy, dx) s[y0 + (dy)][x0 + (dx)]
    const float cfa = S(0, 0);
    float R, G, Bc;
    const bool odd_r = row & 1, odd_c = col & 1;
    if (odd_r == odd_c) {
        // R site (even,even) or B site (odd,odd): G from GR_GB, opposite colour from Rb_BB_Br_RR
        const float g = -0.125f * S(-2, 0) + 0.25f * S(-1, 0) + -0.125f * S(0, -2) + 0.25f * S(0, -1) + 0.5f * cfa +
                        0.25f * S(0, 1) + -0.125f * S(0, 2) + 0.25f * S(1, 0) + -0.125f * S(2, 0);
        const float d = -0.1875f * S(-2, 0) + 0.25f * S(-1, -1) + 0.25f * S(-1, 1) + -0.1875f * S(0, -2) + 0.75f * cfa +
                        -0.1875f * S(0, 2) + 0.25f * S(1, -1) + 0.25f * S(1, 1) + -0.1875f * S(2, 0);
        G = g;
        if (!odd_r) { R = cfa; Bc = d; } else { R = d; Bc = cfa; }
    } else {
        // G sites: row-oriented kernel Rg_RB_Bg_BR and its transpose
        const float hk = 0.0625f * S(-2, 0) + -0.125f * S(-1, -1) + -0.125f * S(-1, 1) + -0.125f * S(0, -2) +
                         0.5f * S(0, -1) + 0.625f * cfa + 0.5f * S(0, 1) + -0.125f * S(0, 2) + -0.125f * S(1, -1) +
                         -0.125f * S(1, 1) + 0.0625f * S(2, 0);
        const float vk = -0.125f * S(-2, 0) + -0.125f * S(-1, -1) + 0.5f * S(-1, 0) + -0.125f * S(-1, 1) +
                         0.0625f * S(0, -2) + 0.625f * cfa + 0.0625f * S(0, 2) + -0.125f * S(1, -1) + 0.5f * S(1, 0) +
                         -0.125f * S(1, 1) + -0.125f * S(2, 0);
        G = cfa;
        if (!odd_r) { R = hk; Bc = vk; }     // green in a red row: R along the row, B along the column
        else        { R = vk; Bc = hk; }     // green in a blue row
    }
#undef S
    const long o = ((long)t * 3) * plane + (long)row * W + col;
    x_rgb[o] = R; x_rgb[o + plane] = G; x_rgb[o + 2 * plane] = Bc;
    if (u) {
        u[o] = R - inv_tau * w[o];
        u[o + plane] = G - inv_tau * w[o + plane];
        u[o + 2 * plane] = Bc - inv_tau * w[o + 2 * plane];
    }
}

extern "C" int sci_malvar2004(const float* x, const float* b, float c_b, const float* w, float inv_tau, float* x_rgb,
                              float* u, int H, int W, int B, void* stream) {
    SCI_REQUIRE(x && x_rgb && H >= 4 && W >= 4 && B > 0 && B <= 65535, "malvar: null/shape");
    SCI_REQUIRE((u == nullptr) == (w == nullptr), "malvar: u and w go together");
    malvar_kernel<<<dim3(sci_ceil_div(W, MV_TW), sci_ceil_div(H, MV_TH), B), dim3(MV_TW, MV_TH), 0, sci_stream(stream)>>>(
        x, b, c_b, w, inv_tau, x_rgb, u, H, W);
    SCI_CHECK_LAUNCH("malvar");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// K3/K5: theta <- clip(RGGB samples of xhat); b += x - theta; w += x_rgb - xhat;
// optional PSNR accumulation of theta against orig.  One thread = one row pair
// segment of 2 columns (a full RGGB quad) per frame.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dual_update_rgb_kernel(const float* __restrict__ xhat, const float* __restrict__ x_rgb,
                                                               float* __restrict__ w, const float* __restrict__ x,
                                                               float* __restrict__ b, float* __restrict__ theta,
                                                               int first_iter, int H, int W,
                                                               const float* __restrict__ orig, double* __restrict__ sse) {
    __shared__ double red[32];
    const int t = blockIdx.z;
    const long plane = (long)H * W;
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2, row = blockIdx.y;
    double err = 0.0;
    if (col < W) {
        const long p = (long)row * W + col;
        const long o = (long)t * 3 * plane + p;
        float2 xh[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xh[c] = *reinterpret_cast<const float2*>(xhat + o + c * plane);
            const float2 xr = *reinterpret_cast<const float2*>(x_rgb + o + c * plane);
            float2 ww = *reinterpret_cast<const float2*>(w + o + c * plane);
            ww.x = ww.x + (xr.x - xh[c].x);
            ww.y = ww.y + (xr.y - xh[c].y);
            *reinterpret_cast<float2*>(w + o + c * plane) = ww;
        }
        // RGGB: even row -> (R, G), odd row -> (G, B)
        const float s0 = (row & 1) ? xh[1].x : xh[0].x;
        const float s1 = (row & 1) ? xh[2].y : xh[1].y;
        const float t0 = fminf(fmaxf(s0, 0.f), 1.f), t1 = fminf(fmaxf(s1, 0.f), 1.f);
        const long q = (long)t * plane + p;
        float2 xv = *reinterpret_cast<const float2*>(x + q);
        if (first_iter) { xv.x = s0; xv.y = s1; }
        float2 bv = *reinterpret_cast<const float2*>(b + q);
        bv.x = bv.x + (xv.x - t0);
        bv.y = bv.y + (xv.y - t1);
        *reinterpret_cast<float2*>(b + q) = bv;
        *reinterpret_cast<float2*>(theta + q) = make_float2(t0, t1);
        if (orig) {
            const float2 og = *reinterpret_cast<const float2*>(orig + q);
            const float d0 = t0 - og.x, d1 = t1 - og.y;
            err = (double)(d0 * d0) + (double)(d1 * d1);
        }
    }
    if (orig) {
        const double s = block_sum(err, red);
        if (threadIdx.x == 0) atomicAdd(sse, s);
    }
}

extern "C" int sci_dual_update_rgb(const float* xhat, const float* x_rgb, float* w, const float* x, float* b,
                                   float* theta, int first_iter, int H, int W, int B, const float* orig, double* sse,
                                   void* stream) {
    SCI_REQUIRE(xhat && x_rgb && w && x && b && theta, "dual_update_rgb: null pointer");
    SCI_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && B > 0 && H <= 65535 && B <= 65535, "dual_update_rgb: shape");
    SCI_REQUIRE(!orig || sse, "dual_update_rgb: orig given without sse");
    dual_update_rgb_kernel<<<dim3(sci_ceil_div(W / 2, 256), H, B), 256, 0, sci_stream(stream)>>>(
        xhat, x_rgb, w, x, b, theta, first_iter, H, W, orig, sse);
    SCI_CHECK_LAUNCH("dual_update_rgb");
    return SCI_OK;
}

// Stage-1 bookkeeping of the deep branches of admm_denoise_bayer_demosaic_pre (dvp:439-503): ONE dual variable,
//   theta = clip(RGGB samples of xhat);  b = b - (x - theta);  PSNR of x (not theta, :507-512).
// first_iter: xall and theta_all are the same tensor at k = 0 (:375-377), so the sampling overwrites x with the UNCLIPPED
// samples before the clip rebinds theta: x is written back here, b receives -(theta_unclipped - theta), the PSNR sees it.
__global__ void __launch_bounds__(256) dual_update_stage1_kernel(const float* __restrict__ xhat, float* __restrict__ x,
                                                                  float* __restrict__ b, float* __restrict__ theta, int first_iter,
                                                                  int H, int W, const float* __restrict__ orig,
                                                                  double* __restrict__ sse) {
    __shared__ double red[32];
    const int t = blockIdx.z;
    const long plane = (long)H * W;
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2, row = blockIdx.y;
    double err = 0.0;
    if (col < W) {
        const long p = (long)row * W + col;
        const long o = (long)t * 3 * plane + p;
        // RGGB: even row -> (R, G), odd row -> (G, B)
        const float2 ca = *reinterpret_cast<const float2*>(xhat + o + ((row & 1) ? 1 : 0) * plane);
        const float2 cb = *reinterpret_cast<const float2*>(xhat + o + ((row & 1) ? 2 : 1) * plane);
        const float s0 = ca.x, s1 = cb.y;
        const float t0 = fminf(fmaxf(s0, 0.f), 1.f), t1 = fminf(fmaxf(s1, 0.f), 1.f);
        const long q = (long)t * plane + p;
        float2 xv = *reinterpret_cast<const float2*>(x + q);
        if (first_iter) {
            xv.x = s0; xv.y = s1;
            *reinterpret_cast<float2*>(x + q) = xv;
        }
        float2 bv = *reinterpret_cast<const float2*>(b + q);
        bv.x = bv.x - (xv.x - t0);
        bv.y = bv.y - (xv.y - t1);
        *reinterpret_cast<float2*>(b + q) = bv;
        *reinterpret_cast<float2*>(theta + q) = make_float2(t0, t1);
        if (orig) {
            const float2 og = *reinterpret_cast<const float2*>(orig + q);
            const float d0 = xv.x - og.x, d1 = xv.y - og.y;
            err = (double)(d0 * d0) + (double)(d1 * d1);
        }
    }
    if (orig) {
        const double s = block_sum(err, red);
        if (threadIdx.x == 0) atomicAdd(sse, s);
    }
}

extern "C" int sci_dual_update_stage1(const float* xhat, float* x, float* b, float* theta, int first_iter, int H, int W, int B,
                                      const float* orig, double* sse, void* stream) {
    SCI_REQUIRE(xhat && x && b && theta, "dual_update_stage1: null pointer");
    SCI_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && B > 0 && H <= 65535 && B <= 65535, "dual_update_stage1: shape");
    SCI_REQUIRE(!orig || sse, "dual_update_stage1: orig given without sse");
    dual_update_stage1_kernel<<<dim3(sci_ceil_div(W / 2, 256), H, B), 256, 0, sci_stream(stream)>>>(xhat, x, b, theta, first_iter,
                                                                                                   H, W, orig, sse);
    SCI_CHECK_LAUNCH("dual_update_stage1");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// Closed-form demosaic update of the `close_form_demosaic` branch (dvp:112-118, 175-182, 224-230), all frames:
//   x_rgb[c] = (rho * x3[c] + b3[c] + tau * xhat[c] + w[c]) / (rho * m[c] + tau)   [clip to [0,1] on the FFDNet branch]
//   u[c]     = x_rgb[c] - inv_tau * w[c]
// where x3 / b3 are the sparse 3-channel images of the Bayer-domain x / b (value at the pixel's CFA channel, 0 elsewhere)
// and m the RGGB mask.  Same operation order as the reference's tensor expression; compiled with --fmad=false.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) closed_form_demosaic_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                                    const float* __restrict__ xhat, const float* __restrict__ w,
                                                                    float rho, float tau, float inv_tau, int clip,
                                                                    float* __restrict__ x_rgb, float* __restrict__ u, int H, int W) {
    const int t = blockIdx.z;
    const long plane = (long)H * W;
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= W) return;
    const long p = (long)row * W + col;
    const int cfa = (row & 1) + (col & 1);                 // RGGB: 0 = R, 1 = G, 2 = B
    const float xv = x[(long)t * plane + p], bv = b[(long)t * plane + p];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long o = ((long)t * 3 + c) * plane + p;
        const float m = (c == cfa) ? 1.f : 0.f;
        const float ww = w[o];
        float num = rho * (m * xv);
        num = num + m * bv;
        num = num + tau * xhat[o];
        num = num + ww;
        float v = num / (rho * m + tau);
        if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
        x_rgb[o] = v;
        u[o] = v - inv_tau * ww;
    }
}

extern "C" int sci_closed_form_demosaic(const float* x, const float* b, const float* xhat, const float* w, float rho, float tau,
                                        float inv_tau, int clip, float* x_rgb, float* u, int H, int W, int B, void* stream) {
    SCI_REQUIRE(x && b && xhat && w && x_rgb && u, "closed_form_demosaic: null pointer");
    SCI_REQUIRE(H > 0 && W > 0 && B > 0 && H <= 65535 && B <= 65535, "closed_form_demosaic: shape");
    closed_form_demosaic_kernel<<<dim3(sci_ceil_div(W, 256), H, B), 256, 0, sci_stream(stream)>>>(x, b, xhat, w, rho, tau, inv_tau,
                                                                                                  clip, x_rgb, u, H, W);
    SCI_CHECK_LAUNCH("closed_form_demosaic");
    return SCI_OK;
}

// Gray-scale variant of the stage-2 bookkeeping (derived FFDNet-gray config): no Bayer sampling,
// theta = clip(xhat); b += x - theta; w += x_pre - xhat; optional PSNR.  One thread = 4 pixels.
__global__ void __launch_bounds__(256) dual_update_gray_kernel(const float* __restrict__ xhat, const float* __restrict__ x_pre,
                                                                float* __restrict__ w, const float* __restrict__ x,
                                                                float* __restrict__ b, float* __restrict__ theta,
                                                                int first_iter, long n, const float* __restrict__ orig,
                                                                double* __restrict__ sse) {
    __shared__ double red[32];
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (i < n) {
        const float xh = xhat[i];
        const float th = fminf(fmaxf(xh, 0.f), 1.f);
        w[i] = w[i] + (x_pre[i] - xh);
        const float xv = first_iter ? xh : x[i];
        b[i] = b[i] + (xv - th);
        theta[i] = th;
        if (orig) { const float d = th - orig[i]; err = (double)(d * d); }
    }
    if (orig) {
        const double s = block_sum(err, red);
        if (threadIdx.x == 0) atomicAdd(sse, s);
    }
}

extern "C" int sci_dual_update_gray(const float* xhat, const float* x_pre, float* w, const float* x, float* b, float* theta,
                                    int first_iter, long n, const float* orig, double* sse, void* stream) {
    SCI_REQUIRE(xhat && x_pre && w && x && b && theta && n > 0, "dual_update_gray: null pointer / size");
    SCI_REQUIRE(!orig || sse, "dual_update_gray: orig given without sse");
    dual_update_gray_kernel<<<sci_ceil_div(n, 256), 256, 0, sci_stream(stream)>>>(xhat, x_pre, w, x, b, theta, first_iter, n,
                                                                                 orig, sse);
    SCI_CHECK_LAUNCH("dual_update_gray");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// RGB <-> Bayer index remaps (bit-exact).
// ---------------------------------------------------------------------------
__global__ void rgb_to_bayer_kernel(const float* __restrict__ rgb, float* __restrict__ mosaic, int H, int W) {
    const int t = blockIdx.z, row = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= W) return;
    const long plane = (long)H * W, p = (long)row * W + col;
    const int c = (row & 1) + (col & 1);                 // (0,0)->R, mixed->G, (1,1)->B
    mosaic[t * plane + p] = rgb[((long)t * 3 + c) * plane + p];
}

__global__ void bayer_to_rgb_sparse_kernel(const float* __restrict__ mosaic, float* __restrict__ rgb, int H, int W) {
    const int t = blockIdx.z, row = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= W) return;
    const long plane = (long)H * W, p = (long)row * W + col;
    const int c = (row & 1) + (col & 1);
    const float v = mosaic[t * plane + p];
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[((long)t * 3 + k) * plane + p] = (k == c) ? v : 0.f;
}

extern "C" int sci_rgb_to_bayer(const float* rgb, float* mosaic, int H, int W, int B, void* stream) {
    SCI_REQUIRE(rgb && mosaic && H > 0 && W > 0 && B > 0 && H <= 65535 && B <= 65535, "rgb_to_bayer");
    rgb_to_bayer_kernel<<<dim3(sci_ceil_div(W, 256), H, B), 256, 0, sci_stream(stream)>>>(rgb, mosaic, H, W);
    SCI_CHECK_LAUNCH("rgb_to_bayer");
    return SCI_OK;
}

extern "C" int sci_bayer_to_rgb_sparse(const float* mosaic, float* rgb, int H, int W, int B, void* stream) {
    SCI_REQUIRE(rgb && mosaic && H > 0 && W > 0 && B > 0 && H <= 65535 && B <= 65535, "bayer_to_rgb_sparse");
    bayer_to_rgb_sparse_kernel<<<dim3(sci_ceil_div(W, 256), H, B), 256, 0, sci_stream(stream)>>>(mosaic, rgb, H, W);
    SCI_CHECK_LAUNCH("bayer_to_rgb_sparse");
    return SCI_OK;
}

// Reference-layout Bayer stack <-> mosaic (API parity with utils/utils_image.py:130-151):
// stack[h][w][B][4]  <->  mosaic[H][W][B].  One thread per element; used only at the API boundary.
__global__ void bayer4_mosaic_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int B,
                                     int to_mosaic) {
    const long n = (long)4 * h * w * B;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // mosaic index
    if (idx >= n) return;
    const int W = 2 * w;
    const int t = (int)(idx % B);
    const long pc = idx / B;
    const int c = (int)(pc % W), r = (int)(pc / W);
    const long sidx = ((((long)(r >> 1) * w + (c >> 1)) * B + t) << 2) + ((r & 1) << 1) + (c & 1);
    if (to_mosaic) dst[idx] = src[sidx]; else dst[sidx] = src[idx];
}

extern "C" int sci_bayer4_to_mosaic(const float* stack, float* mosaic, int h, int w, int B, void* stream) {
    SCI_REQUIRE(stack && mosaic && h > 0 && w > 0 && B > 0, "bayer4_to_mosaic");
    const long n = (long)4 * h * w * B;
    bayer4_mosaic_kernel<<<sci_ceil_div(n, 256), 256, 0, sci_stream(stream)>>>(stack, mosaic, h, w, B, 1);
    SCI_CHECK_LAUNCH("bayer4_to_mosaic");
    return SCI_OK;
}

extern "C" int sci_mosaic_to_bayer4(const float* mosaic, float* stack, int h, int w, int B, void* stream) {
    SCI_REQUIRE(stack && mosaic && h > 0 && w > 0 && B > 0, "mosaic_to_bayer4");
    const long n = (long)4 * h * w * B;
    bayer4_mosaic_kernel<<<sci_ceil_div(n, 256), 256, 0, sci_stream(stream)>>>(mosaic, stack, h, w, B, 0);
    SCI_CHECK_LAUNCH("mosaic_to_bayer4");
    return SCI_OK;
}

// ---------------------------------------------------------------------------
// K11: per-frame sum of squared errors (fp32 terms, fp64 accumulation).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) psnr_accum_kernel(const float* __restrict__ a, const float* __restrict__ orig,
                                                          long npix, double* __restrict__ sse) {
    __shared__ double red[32];
    const int t = blockIdx.y;
    double err = 0.0;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < npix; q += (long)gridDim.x * blockDim.x) {
        const float d = a[t * npix + q] - orig[t * npix + q];
        err += (double)(d * d);
    }
    const double s = block_sum(err, red);
    if (threadIdx.x == 0) atomicAdd(sse + t, s);
}

extern "C" int sci_psnr_accum(const float* a, const float* orig, long npix, int B, double* sse_per_frame, void* stream) {
    SCI_REQUIRE(a && orig && sse_per_frame && npix > 0 && B > 0 && B <= 65535, "psnr_accum");
    const int gx = (int)min((long)SCI_NUM_SMS * 4, (npix + 255) / 256);
    psnr_accum_kernel<<<dim3(gx, B), 256, 0, sci_stream(stream)>>>(a, orig, npix, sse_per_frame);
    SCI_CHECK_LAUNCH("psnr_accum");
    return SCI_OK;
}


// ---------------------------------------------------------------------------
// On-device SSIM for the final per-frame report (dvp:321; skimage.metrics.structural_similarity defaults restated in
// oracle/iqa.py: 7x7 uniform window, sample covariance, K1 = .01, K2 = .03, float64, mean over the image cropped by 3).
// Each thread evaluates S at one pixel of the crop from the 49-tap window sums in fp64 (the crop never touches the
// image border, so the filter's border mode is irrelevant); per-frame sums are block-reduced and added atomically.
// ssim_sum[t] += sum of S over the crop of frame t;  the caller divides by (H-6)*(W-6).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ ref, int H, int W,
                                                    double C1, double C2, double* __restrict__ ssim_sum) {
    __shared__ double red[32];
    const int t = blockIdx.z;
    const long plane = (long)H * W;
    const int col = blockIdx.x * 32 + (threadIdx.x & 31) + 3, row = blockIdx.y * 8 + (threadIdx.x >> 5) + 3;
    double S = 0.0;
    if (col < W - 3 && row < H - 3) {
        const float* pa = a + t * plane + (long)(row - 3) * W + (col - 3);
        const float* pr = ref + t * plane + (long)(row - 3) * W + (col - 3);
        double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
        for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const double x = (double)pr[dy * W + dx], y = (double)pa[dy * W + dx];     // X = reference image, Y = result
                sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
            }
        }
        const double inv = 1.0 / 49.0, cov_norm = 49.0 / 48.0;
        const double ux = sx * inv, uy = sy * inv;
        const double vx = cov_norm * (sxx * inv - ux * ux), vy = cov_norm * (syy * inv - uy * uy);
        const double vxy = cov_norm * (sxy * inv - ux * uy);
        S = ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
    const double sblk = block_sum(S, red);
    if (threadIdx.x == 0) atomicAdd(ssim_sum + t, sblk);
}

extern "C" int sci_ssim_accum(const float* a, const float* ref, int H, int W, int B, double data_range, double* ssim_sum,
                              void* stream) {
    SCI_REQUIRE(a && ref && ssim_sum && H >= 7 && W >= 7 && B > 0 && B <= 65535, "ssim_accum");
    const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
    ssim_kernel<<<dim3(sci_ceil_div(W - 6, 32), sci_ceil_div(H - 6, 8), B), 256, 0, sci_stream(stream)>>>(a, ref, H, W, C1, C2,
                                                                                                          ssim_sum);
    SCI_CHECK_LAUNCH("ssim_accum");
    return SCI_OK;
}


// ---------------------------------------------------------------------------
// Right/bottom reflect padding (torch F.pad(..., mode='reflect'): no edge repeat) and the matching crop for the sequence
// drivers, which pad every frame to a multiple of 4 before calling the network and cut the result back
// (packages/fastdvdnet/fastdvdnet.py:119-141, packages/DDnet/DDnet_test.py:180-196).  planes = any number of [H][W] planes.
// ---------------------------------------------------------------------------
__global__ void reflect_pad_kernel(const float* __restrict__ in, float* __restrict__ out, long planes, int H, int W, int Ho, int Wo) {
    const long total = planes * Ho * Wo;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % Wo), r = (int)((idx / Wo) % Ho);
    const long pl = idx / ((long)Wo * Ho);
    const int rs = r < H ? r : 2 * (H - 1) - r, cs = c < W ? c : 2 * (W - 1) - c;      // Ho <= 2H-1, Wo <= 2W-1 for crop: rs = r
    out[idx] = in[(pl * H + rs) * W + cs];
}

// Right/bottom REPLICATION padding (torch.nn.ReplicationPad2d): KAIR-FFDNet pads odd sizes to even this way before its
// PixelUnShuffle (models/network_ffdnet.py:56-59) and crops afterwards (:68).
__global__ void replicate_pad_kernel(const float* __restrict__ in, float* __restrict__ out, long planes, int H, int W, int Ho, int Wo) {
    const long total = planes * Ho * Wo;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % Wo), r = (int)((idx / Wo) % Ho);
    const long pl = idx / ((long)Wo * Ho);
    out[idx] = in[(pl * H + min(r, H - 1)) * W + min(c, W - 1)];
}

extern "C" int sci_replicate_pad2d(const float* in, float* out, long planes, int H, int W, int Ho, int Wo, void* stream) {
    SCI_REQUIRE(in && out && planes > 0 && H > 0 && W > 0 && Ho >= H && Wo >= W, "replicate_pad2d");
    const long total = planes * Ho * Wo;
    replicate_pad_kernel<<<sci_ceil_div(total, 256), 256, 0, sci_stream(stream)>>>(in, out, planes, H, W, Ho, Wo);
    SCI_CHECK_LAUNCH("replicate_pad2d");
    return SCI_OK;
}

extern "C" int sci_reflect_pad2d(const float* in, float* out, long planes, int H, int W, int Ho, int Wo, void* stream) {
    SCI_REQUIRE(in && out && planes > 0 && H > 0 && W > 0 && Ho >= H && Wo >= W && Ho < 2 * H && Wo < 2 * W, "reflect_pad2d");
    const long total = planes * Ho * Wo;
    reflect_pad_kernel<<<sci_ceil_div(total, 256), 256, 0, sci_stream(stream)>>>(in, out, planes, H, W, Ho, Wo);
    SCI_CHECK_LAUNCH("reflect_pad2d");
    return SCI_OK;
}

__global__ void crop_kernel(const float* __restrict__ in, float* __restrict__ out, long planes, int H, int W, int Hc, int Wc) {
    const long total = planes * Hc * Wc;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % Wc), r = (int)((idx / Wc) % Hc);
    const long pl = idx / ((long)Wc * Hc);
    out[idx] = in[(pl * H + r) * W + c];
}

extern "C" int sci_crop2d(const float* in, float* out, long planes, int H, int W, int Hc, int Wc, void* stream) {
    SCI_REQUIRE(in && out && planes > 0 && Hc > 0 && Wc > 0 && Hc <= H && Wc <= W, "crop2d");
    const long total = planes * Hc * Wc;
    crop_kernel<<<sci_ceil_div(total, 256), 256, 0, sci_stream(stream)>>>(in, out, planes, H, W, Hc, Wc);
    SCI_CHECK_LAUNCH("crop2d");
    return SCI_OK;
}
