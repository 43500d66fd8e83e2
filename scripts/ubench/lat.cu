// Micro-benchmark: dependent-chain latency (cycles per op) of the operations the growth kernel leans on, one warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b, int s)
{
	double x = a + threadIdx.x * 1e-9; float xf = (float)a; int xi = (int)a + threadIdx.x; unsigned xu = xi;
	long long t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < N; i++) {
		if (OP == 0) x = x + b;
		if (OP == 1) x = x * b;
		if (OP == 2) x = x * b + a;          // -fmad=false: mul + add
		if (OP == 3) x = sqrt(x) + b;
		if (OP == 4) x = b / x;
		if (OP == 5) x = floor(x * b) + a;
		if (OP == 6) { xf = (float)x; x = (double)xf + b; }
		if (OP == 7) { xi = (xi * 7 + 3) % s; }
		if (OP == 8) xf = xf * 1.0001f + 0.5f;
		if (OP == 9) xf = rsqrtf(xf) + 1.0f;
		if (OP == 10) { float sn, cs; __sincosf(xf, &sn, &cs); xf = sn + cs; }
		if (OP == 11) xu = __umulhi(xu, 0xD2511F53u) ^ (xu * 0xCD9E8D57u);
		if (OP == 12) xu = __shfl_xor_sync(0xffffffffu, xu, 1) + 1;
		if (OP == 13) xu = __ballot_sync(0xffffffffu, xu & 1) + xu;
		if (OP == 14) xu = __reduce_add_sync(0xffffffffu, xu);
		if (OP == 15) { xi = (int)x; x = (double)xi + b; }
	}
	long long t1 = clock64();
	if (threadIdx.x == 0) cyc[OP] = t1 - t0;
	out[threadIdx.x + 32 * OP] = x + xf + xi + xu;
}
int main()
{
	double* out; long long* cyc; cudaMalloc(&out, 32 * 16 * 8); cudaMallocManaged(&cyc, 16 * 8);
	const char* names[16] = { "dadd", "dmul", "dmul+dadd", "sqrt.rn.f64 (+dadd)", "div.rn.f64", "floor(dmul)+dadd", "f64->f32->f64 (+dadd)", "int mul-add-mod", "ffma", "rsqrtf+fadd", "__sincosf+fadd", "umulhi^mul", "shfl+add", "ballot+add", "redux.add", "f64->i32->f64 (+dadd)" };
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.5, 1.0000001, 977); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(out, cyc, 1.5, 1.0000001, 977); cudaDeviceSynchronize(); printf("%-26s %7.1f cycles/iter\n", names[OP], (double)cyc[OP] / N);
	RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
	return 0;
}
