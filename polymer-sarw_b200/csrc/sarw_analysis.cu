// sarw_analysis.cu — structure analysis of a finished walk on the device.
//
//   rdf_pairs_kernel     scripts/rdf.py:13-20 (rdfPair) and 39-57 (rdfChains): intra- / inter-chain histograms of the
//                        minimum-image distances of all ordered bead pairs, as integer counts
//   chain_shapes_kernel  end-to-end distance^2 and radius of gyration^2 per chain (unwrapped coordinates, sarw.cpp:184)
//
// Results are bit-exact against oracle/sarw_oracle.c (oracle_rdf, oracle_chain_shapes), which is pinned to the reference's own
// rdf.py (tests/golden/rdf_reference.npz). The pair distance is evaluated in fp64 in rdf.py's operation order; a pair is binned by
// comparing its SQUARED distance with per-edge thresholds prepared on the host (the smallest double whose correctly rounded
// square root reaches the edge), so no fp64 square root is taken on the device and the bins are those numpy.histogram gives.
// Beads are sorted by a coarse spatial cell first; tiles of 128 consecutive beads carry bounding boxes and tile pairs that are
// farther apart than the last bin edge are skipped, so a cut-off RDF costs O(N * neighbours) rather than O(N^2).
#include "sarw_device.cuh"
#include "sarw_host.h"
#include <cub/cub.cuh>

namespace sarw {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

constexpr int RDF_TILE = 128;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct RdfBox { double L[3], Linv[3]; int G[3]; };
struct __align__(32) RdfBead { double x, y, z; long long chain; };
struct TileBox { double c[3], h[3]; };   // centre and half width in fractional coordinates

__device__ __forceinline__ double frac_of(double p, double Linv)
{	const double t = p * Linv;
	double v = t - floor(t);
	return v >= 1.0 ? 0.0 : v;
}

__global__ void rdf_keys_kernel(RdfBox b, const double* __restrict__ xyz, int64_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int c[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		int ci = (int)(frac_of(xyz[3 * i + k], b.Linv[k]) * b.G[k]);
		c[k] = ci >= b.G[k] ? b.G[k] - 1 : (ci < 0 ? 0 : ci);
	}
	keys[i] = (uint32_t)((c[2] * b.G[1] + c[1]) * b.G[0] + c[0]);
	idx[i] = (uint32_t)i;
}

// sorted bead records + one bounding box per tile of RDF_TILE consecutive records (block per tile)
__global__ void __launch_bounds__(RDF_TILE) rdf_tiles_kernel(RdfBox b, const double* __restrict__ xyz, const int32_t* __restrict__ chain,
	const uint32_t* __restrict__ order, int64_t n, RdfBead* __restrict__ beads, TileBox* __restrict__ boxes)
{
	const int64_t s = (int64_t)blockIdx.x * RDF_TILE + threadIdx.x;
	double lo[3] = { 2, 2, 2 }, hi[3] = { -1, -1, -1 };
	if (s < n) {
		const uint32_t i = order[s];
		RdfBead r; r.x = xyz[3 * (size_t)i]; r.y = xyz[3 * (size_t)i + 1]; r.z = xyz[3 * (size_t)i + 2]; r.chain = chain[i];
		beads[s] = r;
		const double p[3] = { r.x, r.y, r.z };
#pragma unroll
		for (int k = 0; k < 3; k++) lo[k] = hi[k] = frac_of(p[k], b.Linv[k]);
	}
	__shared__ double slo[3][RDF_TILE / 32], shi[3][RDF_TILE / 32];
#pragma unroll
	for (int k = 0; k < 3; k++) {
#pragma unroll
		for (int d = 16; d; d >>= 1) {
			lo[k] = fmin(lo[k], __shfl_xor_sync(FULL, lo[k], d));
			hi[k] = fmax(hi[k], __shfl_xor_sync(FULL, hi[k], d));
		}
		if ((threadIdx.x & 31) == 0) { slo[k][threadIdx.x >> 5] = lo[k]; shi[k][threadIdx.x >> 5] = hi[k]; }
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		TileBox t;
#pragma unroll
		for (int k = 0; k < 3; k++) {
			double l = slo[k][0], h = shi[k][0];
			for (int w = 1; w < RDF_TILE / 32; w++) { l = fmin(l, slo[k][w]); h = fmax(h, shi[k][w]); }
			t.c[k] = 0.5 * (l + h); t.h[k] = 0.5 * (h - l) + 1e-12;
		}
		boxes[blockIdx.x] = t;
	}
}

// Block = one tile of first members i (one per thread); loops over all tiles of second members j, skipping tiles whose
// bounding box is farther than the last bin edge. Every ORDERED pair (i, j) contributes one count: to `intra` with the
// difference (i - j) when the chains are equal (rdf.py:49-50, self pairs included), otherwise to `inter` with the difference
// (higher chain - lower chain) (rdf.py:51-52: only iChain1 > iChain2 is evaluated and counted twice).
__global__ void __launch_bounds__(RDF_TILE) rdf_pairs_kernel(RdfBox b, const RdfBead* __restrict__ beads, const TileBox* __restrict__ boxes,
	int64_t n, int nTiles, const double* __restrict__ thresh2, int nBins, float rMin, float invDr, double cut2,
	unsigned long long* __restrict__ intra, unsigned long long* __restrict__ inter)
{
	extern __shared__ __align__(32) unsigned char smemRaw[];
	RdfBead* sj = reinterpret_cast<RdfBead*>(smemRaw);                                    // RDF_TILE records
	double* sT = reinterpret_cast<double*>(sj + RDF_TILE);                                // nBins + 1 thresholds
	uint32_t* hist = reinterpret_cast<uint32_t*>(sT + nBins + 1);                         // [2][nBins]
	int* list = reinterpret_cast<int*>(hist + 2 * nBins);                                 // surviving tiles of a batch
	__shared__ int nList;

	for (int k = threadIdx.x; k <= nBins; k += RDF_TILE) sT[k] = thresh2[k];
	for (int k = threadIdx.x; k < 2 * nBins; k += RDF_TILE) hist[k] = 0;
	const int64_t si = (int64_t)blockIdx.x * RDF_TILE + threadIdx.x;
	const bool haveI = si < n;
	RdfBead me; me.x = me.y = me.z = 0; me.chain = 0;
	if (haveI) me = beads[si];
	const TileBox mine = boxes[blockIdx.x];
	__syncthreads();
	const double T0 = sT[0], Tn = sT[nBins];

	for (int base = 0; base < nTiles; base += RDF_TILE) {
		if (threadIdx.x == 0) nList = 0;
		__syncthreads();
		const int t = base + threadIdx.x;
		if (t < nTiles) {
			const TileBox o = boxes[t];
			double lb2 = 0;
#pragma unroll
			for (int k = 0; k < 3; k++) {
				double dc = fabs(mine.c[k] - o.c[k]);
				dc = fmin(dc, 1.0 - dc);
				double gap = dc - (mine.h[k] + o.h[k]);
				gap = gap > 0 ? gap * b.L[k] : 0.0;
				lb2 += gap * gap;
			}
			if (lb2 <= cut2) list[atomicAdd(&nList, 1)] = t;
		}
		__syncthreads();
		const int m = nList;
		for (int li = 0; li < m; li++) {
			const int64_t sjBase = (int64_t)list[li] * RDF_TILE;
			const int cnt = (int)((n - sjBase) < RDF_TILE ? (n - sjBase) : RDF_TILE);
			__syncthreads();
			if (threadIdx.x < cnt) sj[threadIdx.x] = beads[sjBase + threadIdx.x];
			__syncthreads();
			if (!haveI) continue;
#pragma unroll 4
			for (int j = 0; j < cnt; j++) {
				const RdfBead o = sj[j];
				const bool same = me.chain == o.chain;
				const bool flip = me.chain < o.chain;                       // the first member comes from the higher chain
				double d0 = (me.x - o.x) * b.Linv[0], d1 = (me.y - o.y) * b.Linv[1], d2 = (me.z - o.z) * b.Linv[2];   // rdf.py:14
				if (flip) { d0 = -d0; d1 = -d1; d2 = -d2; }                 // (o - me) * Linv, exactly
				d0 -= floor_noxu(0.5 + d0); d1 -= floor_noxu(0.5 + d1); d2 -= floor_noxu(0.5 + d2);                  // rdf.py:15
				d0 *= b.L[0]; d1 *= b.L[1]; d2 *= b.L[2];                                                             // rdf.py:16
				const double r2 = (d0 * d0 + d1 * d1) + d2 * d2;                                                      // rdf.py:17
				if (!(r2 >= T0) || r2 >= Tn) continue;
				int k = (int)((__fsqrt_rn((float)r2) - rMin) * invDr);      // estimate, then settle on the exact thresholds
				k = k < 0 ? 0 : (k >= nBins ? nBins - 1 : k);
				{	// settle on the exact thresholds: the estimate is right or off by one, so the usual cost is two loads and two compares
					const double lo = sT[k], hi = sT[k + 1];
					if (r2 < lo) { do { --k; } while (k > 0 && r2 < sT[k]); }
					else if (r2 >= hi) { do { ++k; } while (k + 1 < nBins && r2 >= sT[k + 1]); }
				}
				atomicAdd(&hist[(same ? 0 : nBins) + k], 1u);
			}
		}
		__syncthreads();
	}
	__syncthreads();
	for (int k = threadIdx.x; k < nBins; k += RDF_TILE) {
		if (hist[k]) atomicAdd(&intra[k], (unsigned long long)hist[k]);
		if (hist[nBins + k]) atomicAdd(&inter[k], (unsigned long long)hist[nBins + k]);
	}
}

size_t rdf_smem_bytes(int nBins)
{	return RDF_TILE * sizeof(RdfBead) + ((size_t)nBins + 1) * sizeof(double) + 2 * (size_t)nBins * sizeof(uint32_t) + RDF_TILE * sizeof(int);
}

// scratch layout: keys, keysOut, idx, idxOut (u32[n] each), beads (RdfBead[n]), boxes (TileBox[nTiles]), cub temp
size_t rdf_scratch_bytes(int64_t n, size_t* cubBytes)
{
	size_t cb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, cb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
	if (cubBytes) *cubBytes = cb;
	const size_t nTiles = ((size_t)n + RDF_TILE - 1) / RDF_TILE;
	auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
	return 4 * up((size_t)n * sizeof(uint32_t)) + up((size_t)n * sizeof(RdfBead)) + up(nTiles * sizeof(TileBox)) + up(cb);
}

cudaError_t launch_rdf(const double box[3], const double* d_xyz, const int32_t* d_chain, int64_t n, const double* d_thresh2, int nBins,
	double rMin, double dr, double lastEdge, void* scratch, unsigned long long* d_intra, unsigned long long* d_inter, cudaStream_t st)
{
	if (n <= 0 || nBins <= 0) return cudaSuccess;
	if (n > 0x7FFFFFF0ll) return cudaErrorInvalidValue;
	RdfBox b;
	for (int k = 0; k < 3; k++) {
		b.L[k] = box[k]; b.Linv[k] = 1. / box[k];
		double edge = lastEdge > 0 ? lastEdge : box[k];
		int G = (int)floor(box[k] / edge);
		b.G[k] = G < 1 ? 1 : (G > 128 ? 128 : G);
	}
	size_t cubBytes = 0; rdf_scratch_bytes(n, &cubBytes);
	const int nTiles = (int)((n + RDF_TILE - 1) / RDF_TILE);
	auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
	unsigned char* p = static_cast<unsigned char*>(scratch);
	uint32_t* keys = (uint32_t*)p; p += up((size_t)n * sizeof(uint32_t));
	uint32_t* keysOut = (uint32_t*)p; p += up((size_t)n * sizeof(uint32_t));
	uint32_t* idx = (uint32_t*)p; p += up((size_t)n * sizeof(uint32_t));
	uint32_t* idxOut = (uint32_t*)p; p += up((size_t)n * sizeof(uint32_t));
	RdfBead* beads = (RdfBead*)p; p += up((size_t)n * sizeof(RdfBead));
	TileBox* boxes = (TileBox*)p; p += up((size_t)nTiles * sizeof(TileBox));
	void* cubTemp = p;
	rdf_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b, d_xyz, n, keys, idx);
	CK(cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, keys, keysOut, idx, idxOut, (int)n, 0, 21, st));
	rdf_tiles_kernel<<<nTiles, RDF_TILE, 0, st>>>(b, d_xyz, d_chain, idxOut, n, beads, boxes);
	const size_t sm = rdf_smem_bytes(nBins);
	CK(cudaFuncSetAttribute(rdf_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
	const double cut = lastEdge * (1 + 1e-9) + 1e-9;
	rdf_pairs_kernel<<<nTiles, RDF_TILE, sm, st>>>(b, beads, boxes, n, nTiles, d_thresh2, nBins, (float)rMin, (float)(1.0 / dr), cut * cut, d_intra, d_inter);
	return cudaGetLastError();
}

// ---- chain shapes: one thread per chain, beads visited in chain order straight from the bead store (slot = 2n + round*n + chain,
// coalesced across the chains of a warp); sums are sequential so the result does not depend on the launch geometry
__global__ void chain_shapes_kernel(const double* __restrict__ lpos, uint32_t idAtomBase, uint32_t nChains, uint32_t rounds,
	double* __restrict__ ree2, double* __restrict__ rg2)
{
	const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nChains) return;
	double sum[3] = { 0, 0, 0 }, first[3] = { 0, 0, 0 }, last[3] = { 0, 0, 0 };
	long long cnt = 0;
	const uint64_t nVisit = 2ull + rounds;
	auto slot_of = [&](uint64_t v) -> uint64_t { return v < 2 ? 2ull * c + v : 2ull * nChains + (v - 2) * nChains + c; };
	for (uint64_t v = 0; v < nVisit; v++) {
		const double* p = lpos + 3 * (size_t)(idAtomBase + slot_of(v));
		const double x = p[0];
		if (__double_as_longlong(x) == -1ll) continue;
		const double y = p[1], z = p[2];
		if (!cnt) { first[0] = x; first[1] = y; first[2] = z; }
		last[0] = x; last[1] = y; last[2] = z;
		sum[0] += x; sum[1] += y; sum[2] += z; cnt++;
	}
	if (!cnt) { ree2[c] = 0; rg2[c] = 0; return; }
	double mean[3], d[3];
#pragma unroll
	for (int k = 0; k < 3; k++) { mean[k] = sum[k] / (double)cnt; d[k] = last[k] - first[k]; }
	ree2[c] = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
	double acc = 0;
	for (uint64_t v = 0; v < nVisit; v++) {
		const double* p = lpos + 3 * (size_t)(idAtomBase + slot_of(v));
		const double x = p[0];
		if (__double_as_longlong(x) == -1ll) continue;
		const double e0 = x - mean[0], e1 = p[1] - mean[1], e2 = p[2] - mean[2];
		acc += (e0 * e0 + e1 * e1) + e2 * e2;
	}
	rg2[c] = acc / (double)cnt;
}

cudaError_t launch_chain_shapes(const Tables& T, uint32_t idAtomBase, uint32_t nChains, uint32_t rounds, double* d_ree2, double* d_rg2, cudaStream_t st)
{
	if (!nChains) return cudaSuccess;
	chain_shapes_kernel<<<(nChains + 127) / 128, 128, 0, st>>>(T.lpos, idAtomBase, nChains, rounds, d_ree2, d_rg2);
	return cudaGetLastError();
}

} // namespace sarw
