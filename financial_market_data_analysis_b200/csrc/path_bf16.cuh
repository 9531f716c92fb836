// path_bf16.cuh - BIGRU_PREC_BF16: bf16 operands on tcgen05 tensor cores (placeholder until the
// tensor-core kernels land; requesting the precision fails loudly, it never falls back to fp32).
#pragma once
#include "common.cuh"

static int bf16_plan_check(const bigru_plan&) {
    bigru_set_error("BIGRU_PREC_BF16 path not built into this library yet");
    return BIGRU_ERR_UNSUPPORTED;
}
static void bf16_workspace(const bigru_plan&, size_t* a, size_t* b) { *a = 0; *b = 0; }
static int forward_bf16(const bigru_plan&, const float*, const float*, const float*, float, int, int, uint64_t, void*,
                        void*, float*, float*, cudaStream_t) { return BIGRU_ERR_UNSUPPORTED; }
static int backward_bf16(const bigru_plan&, const float*, const float*, const float*, float, int, int, uint64_t,
                         const void*, void*, const float*, float*, float*, float*, cudaStream_t) { return BIGRU_ERR_UNSUPPORTED; }
