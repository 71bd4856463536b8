// Experiment (not product): bias of MUFU-based log formulations of the KL uncertainty V_k against fp64.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int K = 4, C = 4;
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// lean log for positive normal x: exponent split + degree-8 minimax-ish (Taylor of log1p on [-0.25,0.5) via atanh form)
__device__ __forceinline__ float lean_log(float x) {
    int i = __float_as_int(x);
    int e = (i - 0x3f3504f3) >> 23;                 // so that m in [sqrt(.5), sqrt(2))
    float m = __int_as_float(i - (e << 23));
    float f = m - 1.0f;
    float s = f / (2.0f + f);                       // replaced below by rcp-based
    float z = s * s;
    float p = fmaf(z, 0.2392828464508056640625f, 0.28518211841583251953125f);
    p = fmaf(p, z, 0.400005877017974853515625f);
    p = fmaf(p, z, 0.666666686534881591796875f);
    float r = fmaf(s * z, p, 2.0f * s);             // 2*atanh(s) = log(m)
    return fmaf((float)e, 0.693147182464599609375f, r);
}
__global__ void k(const float* z, int n, double* out) {
    // out[0..3]: sum V (A: lg2 q & s), (B: lg2 q & p), (C: precise logf), (D: fp64); out[4..7] same for sum E; out[8]: lean
    double a = 0, b = 0, c = 0, d = 0, l = 0;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < n; px += gridDim.x * blockDim.x) {
        float p[K][C], t[K][C], s[K], q[C]; double pd[K][C], qd[C];
        for (int kk = 0; kk < K; ++kk) {
            const float* zz = z + ((size_t)px * K + kk) * C;
            float m = zz[0]; for (int cc = 1; cc < C; ++cc) m = fmaxf(m, zz[cc]);
            float ss = 0; double sd = 0;
            for (int cc = 0; cc < C; ++cc) { t[kk][cc] = (zz[cc] - m) * 1.4426950408889634f; float e = ex2a(t[kk][cc]); p[kk][cc] = e; ss += e;
                pd[kk][cc] = exp((double)zz[cc] - (double)m); sd += pd[kk][cc]; }
            float r = rcpa(ss); s[kk] = ss;
            for (int cc = 0; cc < C; ++cc) { p[kk][cc] *= r; pd[kk][cc] /= sd; }
        }
        for (int cc = 0; cc < C; ++cc) { float acc = p[0][cc]; double ad = pd[0][cc]; for (int kk = 1; kk < K; ++kk) { acc += p[kk][cc]; ad += pd[kk][cc]; } q[cc] = acc * 0.25f; qd[cc] = ad * 0.25; }
        float hA = 0, hC = 0, hL = 0; double hD = 0;
        for (int cc = 0; cc < C; ++cc) { hA += q[cc] * lg2a(q[cc]) * 0.6931471805599453f; hC += q[cc] * logf(q[cc]); hL += q[cc] * lean_log(q[cc]); hD += qd[cc] * log(qd[cc]); }
        for (int kk = 0; kk < K; ++kk) {
            float dA = 0, dB = 0, dC = 0, dL = 0; double dD = 0;
            float ls2 = lg2a(s[kk]), lsC = logf(s[kk]), lsL = lean_log(s[kk]);
            for (int cc = 0; cc < C; ++cc) {
                dA += q[cc] * (t[kk][cc] - ls2) * 0.6931471805599453f;
                dB += q[cc] * lg2a(p[kk][cc]) * 0.6931471805599453f;
                dC += q[cc] * fmaf(t[kk][cc], 0.6931471805599453f, -lsC);
                dL += q[cc] * fmaf(t[kk][cc], 0.6931471805599453f, -lsL);
                dD += qd[cc] * log(pd[kk][cc]);
            }
            a += (double)(hA - dA); b += (double)(hA - dB); c += (double)(hC - dC); d += hD - dD; l += (double)(hL - dL);
        }
    }
    atomicAdd(out + 0, a); atomicAdd(out + 1, b); atomicAdd(out + 2, c); atomicAdd(out + 3, d); atomicAdd(out + 4, l);
}
int main() {
    const int n = 1 << 20;
    for (float scale : {0.02f, 0.05f, 0.2f, 1.0f, 2.0f, 8.0f}) {
        for (float shift : {0.0f, 3.0f}) {
        std::vector<float> h((size_t)n * K * C);
        srand(1234);
        for (auto& v : h) { float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = rand() / (float)RAND_MAX; v = scale * sqrtf(-2 * logf(u1)) * cosf(6.2831853f * u2); }
        for (size_t i = 0; i < h.size(); i += C) h[i] += shift;      // make class 0 dominant: q far from 1/C
        float* dz; double* dout; cudaMalloc(&dz, h.size() * 4); cudaMalloc(&dout, 64); cudaMemset(dout, 0, 64);
        cudaMemcpy(dz, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        k<<<296, 256>>>(dz, n, dout);
        double o[5]; cudaMemcpy(o, dout, 40, cudaMemcpyDeviceToHost);
        printf("scale %.2f shift %.0f: meanV fp64 %.6e | rel err A(lg2 q,s) %.2e  B(lg2 q,p) %.2e  C(logf) %.2e  L(lean) %.2e\n", scale, shift, o[3] / n / K,
               (o[0] - o[3]) / o[3], (o[1] - o[3]) / o[3], (o[2] - o[3]) / o[3], (o[4] - o[3]) / o[3]);
        cudaFree(dz); cudaFree(dout);
        }
    }
    return 0;
}
