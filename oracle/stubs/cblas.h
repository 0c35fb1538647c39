// Stub for OpenBLAS' cblas.h. The reference uses exactly one BLAS call on the path:
// cblas_saxpy(n, 1.0, x, 1, y, 1) in source/kernel/cpu/add_kernel.cpp:13, i.e. y[i] += 1.0f*x[i].
// One fp32 multiply by 1.0 and one fp32 add per element: bit-exact with any BLAS (FMA or not).
#pragma once
static inline void cblas_saxpy(int n, float alpha, const float* x, int incx, float* y, int incy) {
    for (int i = 0; i < n; ++i) y[(long)i * incy] += alpha * x[(long)i * incx];
}
