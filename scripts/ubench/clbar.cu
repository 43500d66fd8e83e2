// Micro-benchmark: latency of barrier.cluster (arrive.release + wait.acquire) vs cluster size and CTA size, with and
// without a DSMEM store / a global store before it.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void k(long long* out, int iters, int mode, float* gbuf)
{
	__shared__ float slot[64];
	cg::cluster_group cl = cg::this_cluster();
	unsigned n = cl.num_blocks(), me = cl.block_rank();
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		if (mode == 1 && threadIdx.x < n) *cl.map_shared_rank(&slot[me], threadIdx.x) = (float)i;
		if (mode == 2 && threadIdx.x == 0) gbuf[blockIdx.x * 32 + (i & 31)] = (float)i;
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}
	long long t1 = clock64();
	if (me == 0 && threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
int main()
{
	long long* out; cudaMallocManaged(&out, 8); float* gbuf; cudaMalloc(&gbuf, 1 << 20);
	cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	for (int threads : { 256, 512 }) for (int mode = 0; mode < 3; mode++) for (int cs : { 1, 2, 4, 7, 8, 13, 16 }) {
		cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(threads);
		cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		cudaError_t e = cudaLaunchKernelEx(&cfg, k, out, 2000, mode, gbuf);
		cudaError_t e2 = cudaDeviceSynchronize();
		printf("threads %d mode %s cluster %2d: %s %lld cycles/barrier\n", threads, mode == 0 ? "bare   " : mode == 1 ? "dsmem  " : "gstore ", cs,
			e != cudaSuccess ? cudaGetErrorString(e) : (e2 != cudaSuccess ? cudaGetErrorString(e2) : "ok"), out[0]);
	}
	return 0;
}
