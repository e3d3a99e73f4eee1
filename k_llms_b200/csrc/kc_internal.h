// kc_internal.h — helpers shared by the translation units of libkllms_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// records the thread-local text kc_last_error() returns and hands `code` back
__attribute__((visibility("hidden"))) int kc_fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

#define KC_CUDA_I(call)                                                                                               \
    do {                                                                                                              \
        cudaError_t e_ = (call);                                                                                      \
        if (e_ != cudaSuccess) return kc_fail(KC_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
