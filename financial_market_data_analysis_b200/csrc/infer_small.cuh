// infer_small.cuh - SURVEY.md 8(f) N5: the live predictor's forward pass (predict.py:165-178) in ONE kernel launch.
// predict.py normalises one window [1, W, F] with the pickled min / max, runs model.forward in eval mode (GRU layers, the
// pooling head of biGRU_model.py:111-137, Linear) and maps the logits through a sigmoid: at the shipped checkpoint's size
// (W = 5, F = 108, H = 8, L = 1) the step-by-step launches of the training path are pure launch latency.  Here one CTA
// per batch row keeps the layer activations in shared memory: thread (direction d, unit u) owns its three gate rows,
// steps are separated by __syncthreads.  Exact expf / tanhf (fp32 parity path).  Meant for small live windows
// (D*H <= 1024 threads, T*max(F, D*H)*8 bytes of shared memory); batches go through bigru_forward.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math.h>

__device__ __forceinline__ float infer_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void infer_window_kernel(const float* __restrict__ params, const float* __restrict__ x, const float* __restrict__ xmin,
                                    const float* __restrict__ xmax, int T, int F, int H, int L, int C, int D,
                                    float* __restrict__ logits, float* __restrict__ probs) {
    extern __shared__ float ism[];
    const int DH = D * H, W = F > DH ? F : DH;
    float* bufA = ism;                         // [T][W]   layer input
    float* bufB = bufA + (size_t)T * W;        // [T][DH]  layer output
    float* hbuf = bufB + (size_t)T * DH;       // [2][DH]  h_{t-1} of every unit (double-buffered)
    float* cat = hbuf + 2 * DH;                // [3H]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int d = tid / H, u = tid % H;
    const bool active = tid < DH;
    // normalised window: (x - min) / (max - min)   (predict.py:170; sql_pytorch_dataloader.py:239)
    for (int i = tid; i < T * F; i += blockDim.x) {
        const int f = i % F;
        float v = x[(int64_t)b * T * F + i];
        if (xmin) v = (v - xmin[f]) / (xmax[f] - xmin[f]);
        bufA[(i / F) * W + f] = v;
    }
    int64_t off = 0;                           // flat-parameter offset of layer l, direction 0 (bigru_b200.h)
    float hlast = 0.f;
    for (int l = 0; l < L; ++l) {
        const int I = l == 0 ? F : DH;
        const int64_t blk = 3LL * H * I + 3LL * H * H + 6LL * H;
        const float* w_ih = params + off + (int64_t)d * blk;
        const float* w_hh = w_ih + 3LL * H * I;
        const float* b_ih = w_hh + 3LL * H * H;
        const float* b_hh = b_ih + 3 * H;
        if (active) hbuf[tid] = 0.f;
        __syncthreads();
        int cur = 0;
        float hown = 0.f;
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? s : T - 1 - s;
            if (active) {
                float ar = b_ih[u], az = b_ih[H + u], an = b_ih[2 * H + u];
                const float* xr = bufA + (size_t)t * W;
                const float* wr = w_ih + (int64_t)u * I, * wz = w_ih + (int64_t)(H + u) * I, * wn = w_ih + (int64_t)(2 * H + u) * I;
                for (int k = 0; k < I; ++k) { const float v = xr[k]; ar = fmaf(wr[k], v, ar); az = fmaf(wz[k], v, az); an = fmaf(wn[k], v, an); }
                float hr = b_hh[u], hz = b_hh[H + u], hn = b_hh[2 * H + u];
                const float* hp = hbuf + cur * DH + d * H;
                const float* vr = w_hh + (int64_t)u * H, * vz = w_hh + (int64_t)(H + u) * H, * vn = w_hh + (int64_t)(2 * H + u) * H;
                for (int k = 0; k < H; ++k) { const float v = hp[k]; hr = fmaf(vr[k], v, hr); hz = fmaf(vz[k], v, hz); hn = fmaf(vn[k], v, hn); }
                const float r = infer_sigmoid(ar + hr), z = infer_sigmoid(az + hz);
                const float n = tanhf(an + r * hn);
                hown = (1.f - z) * n + z * hown;
                bufB[(size_t)t * DH + tid] = hown;
                hbuf[(cur ^ 1) * DH + tid] = hown;
            }
            __syncthreads();
            cur ^= 1;
        }
        hlast = hown;                          // h_n of this (layer, direction)
        // the layer's output is the next layer's input (inter-layer dropout is inactive in eval mode)
        for (int i = tid; i < T * DH; i += blockDim.x) bufA[(i / DH) * W + (i % DH)] = bufB[i];
        __syncthreads();
        off += (int64_t)D * blk;
    }
    // head (biGRU_model.py:111-137): directions summed; last hidden | max over t | mean over t; Linear
    if (active) hbuf[tid] = hlast;
    __syncthreads();
    if (tid < H) {
        float last = hbuf[tid];
        if (D == 2) last += hbuf[H + tid];
        float mx = -INFINITY, sum = 0.f;
        for (int t = 0; t < T; ++t) {
            float sv = bufB[(size_t)t * DH + tid];
            if (D == 2) sv += bufB[(size_t)t * DH + H + tid];
            mx = fmaxf(mx, sv);
            sum += sv;
        }
        cat[tid] = last; cat[H + tid] = mx; cat[2 * H + tid] = sum / (float)T;
    }
    __syncthreads();
    const float* lin_w = params + off;
    const float* lin_b = lin_w + 3LL * H * C;
    for (int c = tid; c < C; c += blockDim.x) {
        float v = lin_b[c];
        for (int k = 0; k < 3 * H; ++k) v = fmaf(cat[k], lin_w[(int64_t)c * 3 * H + k], v);
        logits[(int64_t)b * C + c] = v;
        if (probs) probs[(int64_t)b * C + c] = infer_sigmoid(v);        // predict.py:181
    }
}
