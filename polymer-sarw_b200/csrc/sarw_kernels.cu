// sarw_kernels.cu — the CUDA kernels of libsarw_b200 (sm_100a) and their launch wrappers.
//
//   K1  seed_*            SARW::initiateChain / randomSeed / randomUnitStep     sarw.cpp:290-357, 147-152
//   K2-K4 grow_kernel     SARW::propagateChain / randomConePos / checkCollision sarw.cpp:220-283, 159-214
//                         + PeriodicLookup::find / addPoint                     PeriodicLookup.h:37-75
//                         + main()'s while loop                                 sarw.cpp:67-78
//         <8,false> one GPU, cooperative grid; <8,true> the phases of one melt over several GPUs: replicated tables (phases 1-4)
//         or the box decomposed into z-slabs, one per GPU, each holding its layers of the cell list (phase 5);
//         grow_cluster_kernel (sarw_cluster.cuh) for <= 256 chains: one thread-block cluster, DSMEM, hardware barrier
//   K5  find_batch_kernel PeriodicLookup::find on caller-supplied candidates    PeriodicLookup.h:51-75
//       zbin_*            layer order for large batches (counting sort)
//       add_points_kernel PeriodicLookup::addPoint                              PeriodicLookup.h:37-47
//   K7  download kernels  terminateChain's type flip (sarw.cpp:359-368) + compaction into acceptance order
//   K8  (sarw_analysis.cu) scripts/rdf.py pair histograms, chain shapes
//
// Exact sequential semantics in a parallel round (DESIGN.md "Round protocol"): every chain of the round proposes
// against the beads present at round start and inserts its first passing candidate tentatively (phase A); after a
// grid barrier each chain re-checks its bead against the tentative beads of LOWER-indexed chains (phase V); the
// (rare) chains that conflict are re-resolved one by one in chain order against the true state (clean-up), which
// is exactly what the serial loop of sarw.cpp:251-281 would have produced.
#include "sarw_device.cuh"
#include "sarw_host.h"
#include <cub/device/device_scan.cuh>
#include <cooperative_groups.h>
#include <vector>
#include <cstring>
#include <algorithm>

namespace sarw {

#ifndef SARW_GROW_WARPS
#define SARW_GROW_WARPS 8
#endif
#ifndef SARW_GROW_MINBLOCKS
#define SARW_GROW_MINBLOCKS 2
#endif
constexpr int GROW_WARPS = SARW_GROW_WARPS;   // grid mode: warps per CTA of the cooperative launch (x SARW_GROW_MINBLOCKS CTAs per SM)
constexpr int GROW_NB_CAP = 256;    // grid kernel: staged beads per chain (shared memory, 16 B each)
constexpr int GROW_ARC_CAP = 128;   // grid kernel: blocked arcs per chain
constexpr int GROW_RAW_CELLS = 256; // grid kernel: staged cells per warp = four 4 x 4 x 4 boxes (quad path: four chains per warp at a time)
constexpr int QUAD_NB_CAP = GROW_NB_CAP / 4;   // quad path: staged beads per chain
constexpr uint32_t CHAIN_DEAD = 1u; // ChainState.flags: every torsion angle is permanently blocked
constexpr uint32_t CHAIN_HARD = 2u; // ChainState.flags: none of the first 32 trials passed last round (skip the fast path)
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t DONE_BIT = 0x80000000u;

// ------------------------------------------------------------------ K5: overlap query on caller-supplied candidates
// Large batches are processed in order of the candidates' z cell layer (counting sort by layer: zbin_* kernels below), so that
// candidates running at the same time touch the same two or three layers of the cell table and those lines are served by
// the 126 MB L2 instead of HBM: a 128-byte cell line is then fetched once per batch rather than once per candidate near it.

__device__ __forceinline__ int zbin_of(const Geom& g, const double* __restrict__ xyz, int64_t q)
{
	const double t = xyz[3 * q + 2] * g.Linv[2];
	const int cz = (int)floor((t - floor(t)) * g.sOverL[2]);
	return cz >= g.S[2] ? g.S[2] - 1 : (cz < 0 ? 0 : cz);
}
// Counting sort by layer. Block b owns the candidates [b*chunk, (b+1)*chunk) in both passes. Pass 1 leaves its histogram in
// blockHist[layer * nBlocks + b]; zbin_rows_kernel (one warp per layer) turns every layer's row into exclusive prefixes over the blocks
// and records the layer total; zbin_scan_kernel turns the totals into layer offsets; pass 2 reads its cursors and scatters in ONE pass.
__global__ void __launch_bounds__(512) zbin_count_kernel(Geom g, const double* __restrict__ xyz, int64_t n, int64_t chunk, uint32_t* __restrict__ blockHist)
{
	extern __shared__ uint32_t hist[];
	for (int i = threadIdx.x; i < g.S[2]; i += blockDim.x) hist[i] = 0;
	__syncthreads();
	const int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
	for (int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x) atomicAdd(&hist[zbin_of(g, xyz, q)], 1u);
	__syncthreads();
	for (int i = threadIdx.x; i < g.S[2]; i += blockDim.x) blockHist[(size_t)i * gridDim.x + blockIdx.x] = hist[i];
}
__global__ void __launch_bounds__(256) zbin_rows_kernel(uint32_t* __restrict__ blockHist, uint32_t* __restrict__ layerTotal, int nLayers, int nBlocks)
{
	const int layer = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (layer >= nLayers) return;
	uint32_t* row = blockHist + (size_t)layer * nBlocks;
	uint32_t carry = 0;
	for (int base = 0; base < nBlocks; base += 32) {
		const int b = base + lane;
		const uint32_t v = b < nBlocks ? row[b] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
		if (b < nBlocks) row[b] = carry + incl - v;
		carry += __shfl_sync(FULL, incl, 31);
	}
	if (lane == 0) layerTotal[layer] = carry;
}
__global__ void zbin_scan_kernel(uint32_t* __restrict__ counts, int nb)   // one block: exclusive scan in place (counts -> cursors)
{
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < nb; base += blockDim.x) {
		const int i = base + threadIdx.x;
		uint32_t v = i < nb ? counts[i] : 0u, incl = v;
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		__shared__ uint32_t wsum[32];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
		if (lane == 31) wsum[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0u, wi = w;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(FULL, wi, d); if (lane >= d) wi += y; }
			wsum[lane] = wi - w;
		}
		__syncthreads();
		const uint32_t excl = carry + wsum[warp] + incl - v;
		if (i < nb) counts[i] = excl;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) carry = excl + v;
		__syncthreads();
	}
}
// candidate in processing order: position, ignore index, original index
struct __align__(32) SortedCand { double x, y, z; int64_t ignore; };
__global__ void __launch_bounds__(512) zbin_scatter_kernel(Geom g, const double* __restrict__ xyz, const int64_t* __restrict__ ignore, int64_t n, int64_t chunk,
	const uint32_t* __restrict__ blockHist, const uint32_t* __restrict__ layerBase, SortedCand* __restrict__ rec, uint32_t* __restrict__ perm)
{
	extern __shared__ uint32_t hist[];   // this block's cursor into every layer
	for (int i = threadIdx.x; i < g.S[2]; i += blockDim.x) hist[i] = layerBase[i] + blockHist[(size_t)i * gridDim.x + blockIdx.x];
	__syncthreads();
	const int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
	for (int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x) {
		const double x = xyz[3 * q], y = xyz[3 * q + 1], z = xyz[3 * q + 2];
		const double t = z * g.Linv[2];
		int cz = (int)floor((t - floor(t)) * g.sOverL[2]);
		cz = cz >= g.S[2] ? g.S[2] - 1 : (cz < 0 ? 0 : cz);
		const uint32_t p = atomicAdd(&hist[cz], 1u);
		SortedCand r; r.x = x; r.y = y; r.z = z; r.ignore = ignore ? ignore[q] : -1;
		rec[p] = r;
		perm[p] = (uint32_t)q;
	}
}

// One lane per candidate. The (up to) 8 cells a candidate's sphere touches are fetched with eight independent 256-bit loads (one
// 32-byte cell record = one DRAM sector each) issued back to back, so a warp keeps 256 sectors in flight: the kernel is bound by
// HBM/L2 bandwidth, not by the latency of a chain of dependent loads. Four x-adjacent cells share a 128-byte line, so the two
// cells a sphere touches along x usually arrive with one line fill.
#ifndef SARW_FIND_MINBLOCKS
#define SARW_FIND_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, SARW_FIND_MINBLOCKS) find_batch_kernel(Geom g, Tables T, const double* __restrict__ xyz,
	const int64_t* __restrict__ ignore, int64_t n, uint8_t* __restrict__ hit, const SortedCand* __restrict__ rec, const uint32_t* __restrict__ perm)
{
	// one block per 256 consecutive candidates, no grid-stride loop: blocks are dispatched in order, so the candidates in flight at
	// any time are a narrow window of the processing order (measured: a persistent grid-stride sweep drifts apart and loses the L2 reuse)
	const int64_t qs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (qs < n) {
		int64_t q = qs, ig;
		double pos[3];
		if (rec) {
			const SortedCand r = rec[qs];
			pos[0] = r.x; pos[1] = r.y; pos[2] = r.z; ig = r.ignore; q = (int64_t)perm[qs];
		} else {
			pos[0] = xyz[3 * q]; pos[1] = xyz[3 * q + 1]; pos[2] = xyz[3 * q + 2]; ig = ignore ? ignore[q] : -1;
		}
		double v[3];
		to_frac_noxu(g, pos, v);
		// PeriodicLookup.h:65: ignoreIndex<0 || int(index)!=ignoreIndex
		const uint32_t ignoreId = ig < 0 ? ID_NONE : (uint32_t)ig;
		int c[3][2];
		bool two[3];
		float qo[3][2];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const double s = v[k] * g.sOverL[k];
			double lod;
			const int lo = floor_int_noxu(s - g.reachCells[k], lod);
			two[k] = (s + g.reachCells[k]) >= lod + 1.0;               // reach < 0.5 cell: the sphere touches 1 or 2 cells per dimension
			c[k][0] = wrap_cell(lo, g.S[k]); c[k][1] = wrap_cell(lo + 1, g.S[k]);
			const float d = (float)(lod - s), af = (float)g.a[k];      // one double->float conversion per dimension; errors are inside the band
			qo[k][0] = d * af;
			qo[k][1] = (d + 1.0f) * af;
		}
		uint32_t w[8][CELL_WORDS];
#pragma unroll
		for (int ci = 0; ci < 8; ci++) {
			const int ix = ci & 1, iy = (ci >> 1) & 1, iz = ci >> 2;
			const bool valid = (ix == 0 || two[0]) && (iy == 0 || two[1]) && (iz == 0 || two[2]);
			if (valid) ld_cell(T.cells + cell_index(g, c[0][ix], c[1][iy], c[2][iz]), w[ci]);
			else { w[ci][0] = ID_NONE; w[ci][1] = ID_NONE; w[ci][2] = ID_NONE; w[ci][7] = ID_NONE; }   // empty, not flagged
		}
		bool h = false;
		unsigned flagged = 0;   // bit ci: the cell has beads in the overflow table
#pragma unroll
		for (int ci = 0; ci < 8; ci++) {
			const int ix = ci & 1, iy = (ci >> 1) & 1, iz = ci >> 2;
			flagged |= (cell_flagged(w[ci]) ? 1u : 0u) << ci;
#pragma unroll
			for (int s = 0; s < MAX_CAP; ++s) {
				const uint32_t id = w[ci][s];
				if (h || id == ID_NONE || id == ignoreId) continue;
				h = pair_hits(g, T, v, qo[0][ix] + dec_off(g, cell_off(w[ci], s, 0), 0), qo[1][iy] + dec_off(g, cell_off(w[ci], s, 1), 1),
					qo[2][iz] + dec_off(g, cell_off(w[ci], s, 2), 2), id);
			}
		}
#pragma unroll 1
		while (flagged && !h) {   // a visited cell overflowed (rare): its other beads are in the hash table
			const int ci = __ffs((int)flagged) - 1;
			flagged &= flagged - 1u;
			const bool ix = ci & 1, iy = (ci >> 1) & 1, iz = ci >> 2;
			const float ox = ix ? qo[0][1] : qo[0][0], oy = iy ? qo[1][1] : qo[1][0], oz = iz ? qo[2][1] : qo[2][0];
			ovf_for_cell(T, cell_index(g, ix ? c[0][1] : c[0][0], iy ? c[1][1] : c[1][0], iz ? c[2][1] : c[2][0]),
				[&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
					if (id == ignoreId) return false;
					h = pair_hits(g, T, v, ox + dec_off(g, o0, 0), oy + dec_off(g, o1, 1), oz + dec_off(g, o2, 2), id);
					return h;
				});
		}
		hit[q] = h ? 1 : 0;
	}
}

__global__ void add_points_kernel(Geom g, Tables T, const double* __restrict__ xyz, int64_t n, uint32_t baseId)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double pos[3] = { xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] };
	insert_bead(g, T, pos, baseId + (uint32_t)i);
}

// enumerate beads with id in (idLo, idHi) within thresh of the candidate (clean-up only); lane `sub` of GROUP takes
// cells sub, sub+GROUP, ... including the overflow entries of those cells
template <int GROUP, typename F>
__device__ void enumerate_hits(const Geom& g, const Tables& T, const double v[3], uint32_t idLoExcl, uint32_t idHi, int sub, F&& fn)
{
	int c0[3], n[3];
	float q[3][2];
	for (int k = 0; k < 3; k++) {
		double s = v[k] * g.sOverL[k];
		int lo = (int)floor(s - g.reachCells[k]), hi = (int)floor(s + g.reachCells[k]);
		c0[k] = lo; n[k] = hi - lo + 1;
		q[k][0] = (float)(((double)lo - s) * g.a[k]);
		q[k][1] = (float)(((double)(lo + 1) - s) * g.a[k]);
	}
	const int total = n[0] * n[1] * n[2];
	for (int ci = sub; ci < total; ci += GROUP) {
		const int ix = ci % n[0], rest = ci / n[0];
		const int iy = rest % n[1], iz = rest / n[1];
		const int wz = wrap_cell(c0[2] + iz, g.S[2]);
		const int64_t cidx = cell_index(g, wrap_cell(c0[0] + ix, g.S[0]), wrap_cell(c0[1] + iy, g.S[1]), wz);
		const float qx = q[0][ix], qy = q[1][iy], qz = q[2][iz];
		if (windowed(g) && !in_window(g, wz)) {   // slab context: see probe_cells
			if (g.winStrict) atomicAdd(T.oowCount, 1u);
			ovf_for_cell(T, cidx, [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
				if (id > idLoExcl && id < idHi && pair_hits(g, T, v, qx + dec_off(g, o0, 0), qy + dec_off(g, o1, 1), qz + dec_off(g, o2, 2), id)) fn(id);
				return false;
			});
			continue;
		}
		uint32_t w[CELL_WORDS];
		ld_cell(T.cells + cidx, w);
#pragma unroll
		for (int s = 0; s < MAX_CAP; ++s) {
			const uint32_t id = w[s];
			if (id <= idLoExcl || id >= idHi) continue;
			if (pair_hits(g, T, v, qx + dec_off(g, cell_off(w, s, 0), 0), qy + dec_off(g, cell_off(w, s, 1), 1), qz + dec_off(g, cell_off(w, s, 2), 2), id)) fn(id);
		}
		if (cell_flagged(w))
			ovf_for_cell(T, cidx, [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
				if (id > idLoExcl && id < idHi && pair_hits(g, T, v, qx + dec_off(g, o0, 0), qy + dec_off(g, o1, 1), qz + dec_off(g, o2, 2), id)) fn(id);
				return false;
			});
	}
}

// dirty-list helpers (single thread)
__device__ void dirty_push_unique(volatile uint32_t* list, volatile uint32_t* count, uint32_t chain)
{
	uint32_t n = *count;
	for (uint32_t j = 0; j < n; ++j) if ((list[j] & ~DONE_BIT) == chain) { list[j] = chain; return; }   // already listed: (re-)open it
	list[n] = chain; *count = n + 1;
}
__device__ int dirty_pop_min(volatile uint32_t* list, volatile const uint32_t* count, uint32_t& chain)
{
	uint32_t n = *count, best = ID_NONE; int at = -1;
	for (uint32_t j = 0; j < n; ++j) { uint32_t e = list[j]; if (!(e & DONE_BIT) && e < best) { best = e; at = (int)j; } }
	if (at < 0) return 0;
	list[at] |= DONE_BIT; chain = best;
	return 1;
}

// ------------------------------------------------------------------ K1: seeding (sarw.cpp:290-357)
// six draws of seeding trial t: counter (chain, t, block, 1); see oracle/philox.h for the word -> coordinate map
__device__ __forceinline__ void seed_candidate(const Geom& g, const double* __restrict__ grafts, uint32_t chain, uint32_t trial, double p0[3], double p1[3])
{
	uint4 w0 = philox4x32_10(make_uint4(chain, trial, 0u, 1u), g.key0, g.key1);
	uint4 w1 = philox4x32_10(make_uint4(chain, trial, 1u, 1u), g.key0, g.key1);
	const double sc = 1.0 / 4294967296.0;
	double u[6] = { w0.x * sc, w0.y * sc, w0.z * sc, w0.w * sc, w1.x * sc, w1.y * sc };
	for (int k = 0; k < 3; k++) p0[k] = 0.0 + (g.L[k] - 0.0) * u[2 - k];            // sarw.cpp:336-339 (right-to-left draws)
	if (g.graftLoc == 0 && chain < g.nGraft) { for (int k = 0; k < 3; k++) p0[k] = grafts[3 * (size_t)chain + k]; }   // :342-343
	else if (g.graftLoc == 1 && chain < g.nGraft && g.boundary == 1) p0[2] = (chain % 2) ? 0 : g.L[2];               // :344-345
	else if (g.graftLoc == 2 && chain < g.nGraft && g.boundary == 1) p0[2] = 0.5 * g.L[2];                           // :346-347
	double step[3] = { -.5 + (.5 - -.5) * u[5], -.5 + (.5 - -.5) * u[4], -.5 + (.5 - -.5) * u[3] };   // sarw.cpp:150
	normalize3(step);                                                                                // :151
	for (int k = 0; k < 3; k++) p1[k] = p0[k] + step[k] * 1.54;                                       // :350,352
}

struct SeedArgs {
	Geom g; Tables T;
	ChainState* chains;
	uint32_t* resTrial;
	uint32_t* prev;
	const double* grafts;
	Control* ctl;
	uint32_t* dirtyList;
};

// does the seed pair pass checkCollision (sarw.cpp:353,193-204) against beads with id < idHi?
__device__ __forceinline__ bool seed_pair_ok(const Geom& g, const Tables& T, const double p0[3], const double p1[3], uint32_t idHi)
{
	if (walls_reject(g, p0) || walls_reject(g, p1)) return false;
	double v[3];
	to_frac(g, p0, v);
	if (probe_cells<1>(g, T, v, 0u, idHi, g.seedIgnoreId, 0)) return false;
	to_frac(g, p1, v);
	if (probe_cells<1>(g, T, v, 0u, idHi, g.seedIgnoreId, 0)) return false;
	return true;
}

__global__ void __launch_bounds__(256) seed_propose_kernel(SeedArgs A)
{
	const Geom& g = A.g;
	const int lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
	const uint32_t usable = g.maxTrials > 0 ? (uint32_t)g.maxTrials - 1u : 0u;   // sarw.cpp:355: success on trial maxTrials counts as failure
	for (uint32_t i = gw; i < g.nChains; i += nW) {
		uint32_t accepted = 0;
		double p0[3], p1[3];
		for (uint32_t base = 0; base < usable && !accepted; base += 32) {
			uint32_t t = base + lane + 1;
			bool ok = t <= usable;
			if (ok) { seed_candidate(g, A.grafts, i, t, p0, p1); ok = seed_pair_ok(g, A.T, p0, p1, g.idAtomBase); }
			unsigned m = __ballot_sync(FULL, ok);
			if (m) {
				int first = __ffs(m) - 1;
				accepted = base + first + 1;
				if (lane == first) {
					insert_bead(g, A.T, p0, g.idAtomBase + 2 * i);
					insert_bead(g, A.T, p1, g.idAtomBase + 2 * i + 1);
				}
			}
		}
		if (lane == 0) {
			A.resTrial[i] = accepted ? accepted : (uint32_t)g.maxTrials + 1u;
			if (!accepted) atomicMin(&A.ctl->seedFailChain, i);
		}
	}
}

__global__ void __launch_bounds__(256) seed_verify_kernel(SeedArgs A)
{
	const Geom& g = A.g;
	const int sub = threadIdx.x & 7, grp = (threadIdx.x & 31) >> 3;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t base = gw * 4; base < g.nChains; base += nW * 4) {
		uint32_t i = base + grp;
		bool h = false;
		if (i < g.nChains && __ldcg(&A.resTrial[i]) <= (uint32_t)g.maxTrials && i > 0) {
			for (int b = 0; b < 2 && !h; b++) {
				const double* lp = A.T.lpos + 3 * (size_t)(g.idAtomBase + 2 * i + b);
				double pos[3] = { __ldcg(lp), __ldcg(lp + 1), __ldcg(lp + 2) }, v[3];
				to_frac(g, pos, v);
				h = probe_cells<8>(g, A.T, v, g.idAtomBase, g.idAtomBase + 2 * i, ID_NONE, sub);
			}
		}
		unsigned bal = __ballot_sync(FULL, h);
		if (sub == 0 && ((bal >> (grp * 8)) & 0xFFu)) {
			uint32_t at = atomicAdd(&A.ctl->seedDirtyCount, 1u);
			A.dirtyList[at] = i;
		}
	}
}

// single thread: re-resolve conflicting seeds in chain order; then discard everything from the first failing chain on
__global__ void seed_cleanup_kernel(SeedArgs A)
{
	const Geom& g = A.g;
	const uint32_t T_ = (uint32_t)g.maxTrials, usable = T_ > 0 ? T_ - 1u : 0u;
	uint32_t chain;
	while (dirty_pop_min(A.dirtyList, &A.ctl->seedDirtyCount, chain)) {
		if (chain >= A.ctl->seedFailChain) break;
		const uint32_t id0 = g.idAtomBase + 2 * chain, id1 = id0 + 1;
		uint32_t t0 = A.resTrial[chain];
		remove_bead(g, A.T, id0); remove_bead(g, A.T, id1);
		uint32_t found = 0; double p0[3], p1[3];
		for (uint32_t t = t0; t <= usable; ++t) {
			seed_candidate(g, A.grafts, chain, t, p0, p1);
			if (seed_pair_ok(g, A.T, p0, p1, id0)) { found = t; break; }
		}
		A.ctl->conflicts++;
		if (!found) { A.resTrial[chain] = T_ + 1u; if (chain < A.ctl->seedFailChain) A.ctl->seedFailChain = chain; continue; }
		A.resTrial[chain] = found;
		insert_bead(g, A.T, p0, id0); insert_bead(g, A.T, p1, id1);
		__threadfence();
		const uint32_t idEnd = g.idAtomBase + 2 * g.nChains;
		auto mark = [&](uint32_t id) { dirty_push_unique(A.dirtyList, &A.ctl->seedDirtyCount, (id - g.idAtomBase) >> 1); };
		double v[3];
		to_frac(g, p0, v); enumerate_hits<1>(g, A.T, v, id1, idEnd, 0, mark);
		to_frac(g, p1, v); enumerate_hits<1>(g, A.T, v, id1, idEnd, 0, mark);
	}
	const uint32_t fail = A.ctl->seedFailChain;
	if (fail != ID_NONE)   // sarw.cpp:313-314: initiateChain returns at the first failing chain; later chains are never seeded
		for (uint32_t i = fail; i < g.nChains; ++i)
			if (A.resTrial[i] <= T_) { remove_bead(g, A.T, g.idAtomBase + 2 * i); remove_bead(g, A.T, g.idAtomBase + 2 * i + 1); A.resTrial[i] = T_ + 1u; }
}

__global__ void seed_commit_kernel(SeedArgs A)
{
	const Geom& g = A.g;
	const uint32_t fail = A.ctl->seedFailChain;
	const uint32_t nSeeded = fail == ID_NONE ? g.nChains : fail;
	unsigned long long myTrials = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.nChains; i += gridDim.x * blockDim.x) {
		ChainState cs;
		memset(&cs, 0, sizeof cs);
		cs.lastId = cs.penId = ID_NONE;
		if (i < nSeeded) {
			const uint32_t id0 = g.idAtomBase + 2 * i, id1 = id0 + 1;
			for (int k = 0; k < 3; k++) { cs.pen[k] = A.T.lpos[3 * (size_t)id0 + k]; cs.last[k] = A.T.lpos[3 * (size_t)id1 + k]; }
			cs.penId = id0; cs.lastId = id1; cs.length = 2;   // sarw.cpp:306-310
			A.prev[2 * i] = ID_NONE; A.prev[2 * i + 1] = id0;
			myTrials += A.resTrial[i];
		} else if (i == nSeeded) myTrials += (unsigned)g.maxTrials;
		A.chains[i] = cs;
		A.resTrial[i] = 0;
	}
	if (myTrials) atomicAdd(&A.ctl->seedTrials, myTrials);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		A.ctl->nBeads = 2ull * nSeeded;
		A.ctl->nSeeded = nSeeded;
		A.ctl->initiated = fail == ID_NONE ? 1u : 0u;
	}
}

// ------------------------------------------------------------------ K2-K4: persistent growth kernel
// tentative bead of one chain as exchanged between ranks (all-gather once per round)
struct __align__(8) MrRecord { double x, y, z; uint32_t trial; uint32_t pad; };
static_assert(sizeof(MrRecord) == 32, "MrRecord is 32 bytes");

struct GrowArgs {
	Geom g; Tables T;
	ChainState* chains;
	uint32_t* resTrial;     // per chain: 0 nothing pending, 1..T accepted trial of the running round, T+1 failed
	uint32_t* origTrial;    // per chain: the trial phase A stopped at this round (conflict resolution restarts there)
	double* resPos;         // per chain: tentative / accepted bead of the running round
	uint32_t* prev;         // per atom slot: lookup index of the previous bead of its chain (bond partner, sarw.cpp:272)
	Control* ctl;
	uint32_t* roundAcc;     // per round: beads accepted
	uint32_t* roundProg;    // per round: chains accepted before their last trial (sarw.cpp:277)
	uint32_t* dirtyCount;   // per round
	uint32_t* dirtyList;
	uint32_t roundEnd;      // run rounds [ctl->round, roundEnd)
	// multi-rank mode (one melt over several GPUs, replicated state, partitioned proposals)
	uint32_t mrRank, mrRanks;     // chain i is proposed by rank i % mrRanks
	uint32_t mrPhase;             // 0 whole rounds (single GPU); 1 commit + propose one round; 2 apply remote + verify + clean-up; 3 commit only
	uint32_t mrK;                 // records per rank = ceil(nChains / mrRanks)
	MrRecord* mrLocal;            // phase 1 output: this rank's records, chain i at [i / mrRanks]
	const MrRecord* mrAll;        // phase 2 input: all ranks' records, chain i at [(i % mrRanks) * mrK + i / mrRanks]
	// phase 4: whole loop in one launch per GPU; records are stored straight into every peer's buffer over NVLink and the
	// ranks meet on flags in peer memory (no host, no NCCL in the loop)
	MrRecord* peerRec[8];         // [r]: rank r's record buffer, 2 parities x mrRanks x mrK records (own buffer at [mrRank])
	uint32_t* peerFlags[8];       // [r]: rank r's arrival flags: 2 parities x 8 words for the records, then 2 x 8 for the conflict lists
	uint32_t* peerDirty[8];       // [r]: rank r's conflict-list inbox, 2 parities x mrRanks x (1 + mrK) words (count, chains...)
	uint32_t mrSplitVerify;       // phase 4: each rank verifies only the chains it proposed and the conflict lists are exchanged
	// phase 5: slab decomposition. Every rank holds the cell layers of its z-slab plus a halo; a chain is proposed,
	// verified and re-resolved by the rank whose slab holds its tip; records (one slot per chain), conflict lists and re-resolution
	// results travel through peer memory; chain states and bead positions (lpos) stay replicated.
	uint32_t* ownList;            // chains whose tip lies in this rank's slab this round (any order)
	uint8_t* ownerOf;             // per chain: the rank that owns it this round
	MrRecord* cleanRes;           // re-resolutions this rank made in the running clean-up iteration: {x, y, z, trial, pad = chain}
	uint32_t* cleanPush;          // chains those re-resolutions put in conflict
	uint32_t cleanCap;            // capacity of both
	unsigned char* peerClean[8];  // [r]: rank r's clean-up inbox, 2 parities x mrRanks messages of clean_msg_bytes(cleanCap)
	volatile unsigned* trace;     // diagnostics (SARW_TRACE_MAPPED=1): mapped host memory, block 0 / thread 0 stores round << 8 | stage before every phase
	alignas(64) unsigned char tmap[128];   // CUtensorMap of the cell table (valid when g.useTmap); read by TMA straight from the kernel parameters
};

__device__ __forceinline__ uint32_t round_id_base(const Geom& g, uint32_t round)
{	return g.idAtomBase + 2u * g.nChains + round * g.nChains;
}

// chain range of a round, sarw.cpp:230-249
__device__ __forceinline__ void round_range(const Geom& g, unsigned long long nBeads, uint32_t& start, uint32_t& end)
{
	start = 0; end = g.nChains;
	if (g.growthOrder == 1 && g.nGraft > 0) {
		if ((long long)nBeads < g.nGraftAtoms) end = g.nGraft; else start = g.nGraft;
	}
}

// exact evaluation of one trial by the whole warp (reference arithmetic; neighbours split across lanes). The neighbour list
// is complete: region_consume appended the overflow-table beads of the staged cells.
__device__ __forceinline__ bool eval_trial_coop(const Geom& g, const Tables& T, const Frame& f, const ChainState& cs, double u01,
	const float4* nb, int nnb, uint32_t idRound, int lane, double cand[3])
{
	cone_candidate(g, f, cs.last, u01, cand);
	if (growth_rules_reject(g, cand, cs.last)) return false;
	double v[3];
	to_frac(g, cand, v);
	bool hit = false;
	if (nnb >= 0) {
		const float rx = (float)(cand[0] - cs.last[0]), ry = (float)(cand[1] - cs.last[1]), rz = (float)(cand[2] - cs.last[2]);
		for (int j = lane; j < nnb && !hit; j += 32) {
			const float4 b = nb[j];
			const float dx = rx - b.x, dy = ry - b.y, dz = rz - b.z;
			const float d2 = dx * dx + dy * dy + dz * dz;
			if (d2 < g.lo2) hit = true;
			else if (d2 <= g.hi2) hit = exact_pair(g, v, T.lpos + 3 * (size_t)__float_as_uint(b.w));
		}
	} else {   // region too crowded for the staging buffer: query the cell list directly (8 cells over 8 lanes)
		if (lane < 8) hit = probe_cells<8>(g, T, v, 0u, idRound, cs.lastId, lane);
	}
	return !__any_sync(FULL, hit);
}

// publish one chain's record: phases 1/2 into the local staging buffer, phase 4 into every rank's buffer (peer stores)
__device__ __forceinline__ void mr_publish(const GrowArgs& A, uint32_t i, uint32_t round, const MrRecord& r)
{
	if (A.mrPhase == 5u) {   // slab mode: one slot per chain (ownership moves with the tip)
		const size_t at = (size_t)(round & 1u) * A.g.nChains + i;
		for (uint32_t p = 0; p < A.mrRanks; ++p) A.peerRec[p][at] = r;
	} else if (A.mrPhase == 4u) {
		const size_t at = (size_t)(round & 1u) * A.mrRanks * A.mrK + (size_t)A.mrRank * A.mrK + i / A.mrRanks;
		for (uint32_t p = 0; p < A.mrRanks; ++p) A.peerRec[p][at] = r;
	} else A.mrLocal[i / A.mrRanks] = r;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{	uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// one thread per GPU: tell every peer that this rank's records of `round` are complete, wait until all peers said so.
// Returns false after ~3 s without progress (a peer died): the caller aborts the launch instead of hanging the GPU.
__device__ bool mr_peer_meet_tag(const GrowArgs& A, uint32_t par, uint32_t tag, uint32_t which)
{
	const uint32_t off = which * 16u + par * 8u;
	__threadfence_system();
	for (uint32_t p = 0; p < A.mrRanks; ++p) if (p != A.mrRank) st_release_sys(A.peerFlags[p] + off + A.mrRank, tag);
	const long long t0 = clock64();
	for (uint32_t p = 0; p < A.mrRanks; ++p) {
		if (p == A.mrRank) continue;
		while (ld_acquire_sys(A.peerFlags[A.mrRank] + off + p) != tag)
			if (clock64() - t0 > 6000000000ll) return false;
	}
	return true;
}
__device__ __forceinline__ bool mr_peer_meet(const GrowArgs& A, uint32_t round, uint32_t which = 0u) { return mr_peer_meet_tag(A, round & 1u, round + 1u, which); }

// apply the result of the previous round to the chain state (sarw.cpp:270-273); all lanes compute, lane 0 stores
__device__ __forceinline__ void commit_pending(const GrowArgs& A, uint32_t i, uint32_t pendingRound, ChainState& cs, int lane)
{
	const Geom& g = A.g;
	uint32_t t = __ldcg(&A.resTrial[i]);
	if (t >= 1u && t <= (uint32_t)g.maxTrials) {
		const uint32_t id = round_id_base(g, pendingRound) + i;
		for (int k = 0; k < 3; k++) { cs.pen[k] = cs.last[k]; cs.last[k] = __ldcg(&A.resPos[3 * (size_t)i + k]); }
		if (lane == 0) A.prev[id - g.idAtomBase] = cs.lastId;
		cs.penId = cs.lastId; cs.lastId = id; cs.length += 1;
		if (lane == 0) A.chains[i] = cs;
	}
	if (t != 0 && lane == 0) A.resTrial[i] = 0;
}

// the same commit by ONE thread (multi-rank: chains proposed by other ranks are committed thread-per-chain, fully parallel)
__device__ __forceinline__ void commit_pending_thread(const GrowArgs& A, uint32_t i, uint32_t pendingRound)
{
	const Geom& g = A.g;
	const uint32_t t = __ldcg(&A.resTrial[i]);
	if (t >= 1u && t <= (uint32_t)g.maxTrials) {
		const uint32_t id = round_id_base(g, pendingRound) + i;
		ChainState* p = A.chains + i;
		const uint32_t oldLast = __ldcg(&p->lastId);
		for (int k = 0; k < 3; k++) { p->pen[k] = __ldcg(&p->last[k]); }
		for (int k = 0; k < 3; k++) { p->last[k] = __ldcg(&A.resPos[3 * (size_t)i + k]); }
		A.prev[id - g.idAtomBase] = oldLast;
		p->penId = oldLast; p->lastId = id; p->length = __ldcg(&p->length) + 1u;
	}
	if (t != 0u) A.resTrial[i] = 0;
}

__device__ __forceinline__ ChainState load_chain(const ChainState* p)
{
	ChainState cs;
	const double2* q = reinterpret_cast<const double2*>(p);   // 64 B = 4 x 16 B, L2 (cache-global) loads
	double2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3);
	cs.last[0] = a.x; cs.last[1] = a.y; cs.last[2] = b.x; cs.pen[0] = b.y; cs.pen[1] = c.x; cs.pen[2] = c.y;
	unsigned long long u = (unsigned long long)__double_as_longlong(d.x), w = (unsigned long long)__double_as_longlong(d.y);
	cs.lastId = (uint32_t)u; cs.penId = (uint32_t)(u >> 32); cs.length = (uint32_t)w; cs.flags = (uint32_t)(w >> 32);
	return cs;
}

// Prefetch of everything phase A needs to open a chain: one 32-bit word per lane (lanes 0-15 the chain state, 16-21 the position
// the last round settled on, 22 its trial), issued while the previous chain of this warp is processed, so the two dependent
// round trips to L2/HBM (state + trial, then position: 2.7k cycles per chain, measured) disappear from the per-chain latency.
__device__ __forceinline__ uint32_t chain_prefetch(const GrowArgs& A, uint32_t i, int lane)
{
	const uint32_t* p = lane < 16 ? reinterpret_cast<const uint32_t*>(A.chains + i) + lane
		: lane < 22 ? reinterpret_cast<const uint32_t*>(A.resPos + 3 * (size_t)i) + (lane - 16)
		: reinterpret_cast<const uint32_t*>(A.resTrial + i);
	return lane < 23 ? __ldcg(p) : 0u;
}
__device__ __forceinline__ double pf_double(uint32_t pf, int w)
{	return __hiloint2double((int)__shfl_sync(FULL, pf, w + 1), (int)__shfl_sync(FULL, pf, w));
}
// state of chain i from its prefetched words, with the pending result of `pendingRound` committed (commit_pending's effect)
__device__ __forceinline__ ChainState chain_open(const GrowArgs& A, uint32_t i, uint32_t pf, bool pendingValid, uint32_t pendingRound, int lane)
{
	const Geom& g = A.g;
	ChainState cs;
#pragma unroll
	for (int k = 0; k < 3; k++) { cs.last[k] = pf_double(pf, 2 * k); cs.pen[k] = pf_double(pf, 6 + 2 * k); }
	cs.lastId = __shfl_sync(FULL, pf, 12); cs.penId = __shfl_sync(FULL, pf, 13); cs.length = __shfl_sync(FULL, pf, 14); cs.flags = __shfl_sync(FULL, pf, 15);
	if (pendingValid) {
		const uint32_t t = __shfl_sync(FULL, pf, 22);
		if (t >= 1u && t <= (uint32_t)g.maxTrials) {
			const uint32_t id = round_id_base(g, pendingRound) + i;
#pragma unroll
			for (int k = 0; k < 3; k++) { cs.pen[k] = cs.last[k]; cs.last[k] = pf_double(pf, 16 + 2 * k); }
			if (lane == 0) A.prev[id - g.idAtomBase] = cs.lastId;
			cs.penId = cs.lastId; cs.lastId = id; cs.length += 1;
			if (lane == 0) A.chains[i] = cs;
		}
		if (t != 0 && lane == 0) A.resTrial[i] = 0;
	}
	return cs;
}

// ------------------------------------------------------------------ same-round conflict resolution
// re-resolve ONE conflicting chain with the whole warp: drop its current bead, search from the trial phase A stopped at
// against the beads of the previous rounds plus this round's beads of LOWER chains, insert the new bead, report the
// higher chains of this round that now collide with it. Lanes 8g..8g+7 evaluate trial t0+g (one cell each).
template <typename Push>
__device__ void resolve_dirty_chain(const GrowArgs& A, uint32_t round, uint32_t chain, int lane, Push&& push)
{
	const Geom& g = A.g;
	const uint32_t T_ = (uint32_t)g.maxTrials;
	const uint32_t idRound = round_id_base(g, round), idEnd = idRound + g.nChains;
	const int sub = lane & 7, grp = lane >> 3;
	const uint32_t myId = idRound + chain;
	const ChainState cs = load_chain(A.chains + chain);
	Frame f; cone_frame(cs.last, cs.pen, f);
	const uint32_t t0 = __ldcg(&A.origTrial[chain]);       // where phase A stopped: trials before it fail against old beads alone
	const uint32_t cur = __ldcg(&A.resTrial[chain]);        // what is booked (and inserted) right now
	if (lane == 0 && cur <= T_) remove_bead(g, A.T, myId);
	__syncwarp();
	uint32_t found = 0; double cand[3] = { 0, 0, 0 }, v[3] = { 0, 0, 0 };
	for (uint32_t tb = t0; tb <= T_ && !found; tb += 4) {
		const uint32_t t = tb + grp;
		bool rejected = t > T_, hit = false;
		if (!rejected) {
			cone_candidate(g, f, cs.last, u_grow(g, chain, round, t), cand);
			rejected = growth_rules_reject(g, cand, cs.last);
			if (!rejected) { to_frac(g, cand, v); hit = probe_cells<8>(g, A.T, v, 0u, myId, cs.lastId, sub); }
		}
		const unsigned hm = __ballot_sync(FULL, hit || rejected);
		for (int q = 0; q < 4 && !found; ++q)
			if (((hm >> (8 * q)) & 0xFFu) == 0u) {   // first group (lowest trial) whose 8 lanes all report "free"
				found = tb + q;
				for (int k2 = 0; k2 < 3; k2++) { cand[k2] = __shfl_sync(FULL, cand[k2], 8 * q); v[k2] = __shfl_sync(FULL, v[k2], 8 * q); }
			}
	}
	if (lane == 0) {
		atomicAdd(&A.ctl->conflicts, 1ull);
		const long long before = cur <= T_ ? cur : T_, after = found ? found : T_;
		atomicAdd(&A.ctl->trials, (unsigned long long)(after - before));
		atomicAdd(&A.ctl->queries, (unsigned long long)(after - (long long)t0 + 1));
		const int hadBead = cur <= T_, hasBead = found != 0u, hadProg = cur < T_, hasProg = found != 0u && found < T_;
		if (hadBead != hasBead) atomicAdd(&A.roundAcc[round], (uint32_t)(hasBead - hadBead));
		if (hadProg != hasProg) atomicAdd(&A.roundProg[round], (uint32_t)(hasProg - hadProg));
		A.resTrial[chain] = found ? found : T_ + 1u;
		if (found) {
			for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)chain + k2] = cand[k2];
			insert_bead(g, A.T, cand, myId);
			__threadfence();
		}
	}
	__syncwarp();
	if (found) {   // higher-indexed chains of this round whose tentative bead now collides with the new position
		uint32_t hits[8]; int nh = 0;
		if (lane < 8) enumerate_hits<8>(g, A.T, v, myId, idEnd, lane, [&](uint32_t id) { if (nh < 8) hits[nh] = id; nh++; });
		unsigned pending = __ballot_sync(FULL, nh > 0);
		unsigned spilled = __ballot_sync(FULL, nh > 8);   // more hits than the register list holds (tentative beads of one round may pile up)
		while (pending) {
			const int src = __ffs(pending) - 1;
			const int cnt = min(__shfl_sync(FULL, nh, src), 8);
			for (int q = 0; q < cnt; ++q) {
				uint32_t id = 0;
#pragma unroll
				for (int z = 0; z < 8; ++z) if (z == q) id = hits[z];
				id = __shfl_sync(FULL, id, src);
				if (lane == 0) push(id - idRound);
			}
			pending &= pending - 1;
		}
		while (spilled) {   // the lane enumerates again and reports what did not fit, one lane at a time (push is not re-entrant)
			const int src = __ffs(spilled) - 1;
			if (lane == src) { int k = 0; enumerate_hits<8>(g, A.T, v, myId, idEnd, lane, [&](uint32_t id) { if (k++ >= 8) push(id - idRound); }); }
			__syncwarp();
			spilled &= spilled - 1;
		}
		__syncwarp();
	}
}

// squared minimum-image distance between the (round-start) tips of two chains
__device__ __forceinline__ double tip_dist2(const GrowArgs& A, uint32_t a, uint32_t b)
{
	double d2 = 0;
	for (int k = 0; k < 3; k++) {
		double d = __ldcg(&A.chains[a].last[k]) - __ldcg(&A.chains[b].last[k]);
		d -= A.g.L[k] * rint(d * A.g.Linv[k]);
		d2 += d * d;
	}
	return d2;
}

// radius beyond which two conflicting chains cannot influence each other's re-resolution nor touch the same cells:
// two bonds + minDist for the geometry, four cell edges so that their beads and the cells they probe never overlap
__device__ __forceinline__ double independence_radius(const Geom& g)
{	const double amax = fmax(g.a[0], fmax(g.a[1], g.a[2]));
	return 2 * 1.54 + g.thresh + 4 * amax;
}

// PARALLEL part (every warp of the grid): a conflicting chain with no LOWER conflicting chain within the independence
// radius depends only on final beads, so all such chains are re-resolved at once, one per warp. Chains they newly put in
// conflict go to the second list and are handled, with the dependent ones, by the ordered pass below.
__device__ void cleanup_parallel(const GrowArgs& A, uint32_t round, uint32_t gw, uint32_t nW, int lane)
{
	const uint32_t nd = __ldcg(&A.dirtyCount[round]);
	const double R2 = independence_radius(A.g) * independence_radius(A.g);
	for (uint32_t e = gw; e < nd; e += nW) {
		const uint32_t chain = __ldcg(&A.dirtyList[e]);
		bool dep = false;
		for (uint32_t j = lane; j < nd && !dep; j += 32) {
			const uint32_t other = __ldcg(&A.dirtyList[j]) & ~DONE_BIT;
			if (other < chain && tip_dist2(A, chain, other) <= R2) dep = true;
		}
		if (__any_sync(FULL, dep)) { if (lane == 0) A.ctl->needOrdered[round & 1u] = 1u; continue; }
		resolve_dirty_chain(A, round, chain, lane, [&](uint32_t c) { const uint32_t at = atomicAdd(&A.ctl->dirty2Count, 1u); A.dirtyList[A.g.nChains + at] = c;
			A.ctl->needOrdered[round & 1u] = 1u; });
		if (lane == 0) A.dirtyList[e] = chain | DONE_BIT;
		__syncwarp();
	}
}

// ORDERED part (one warp): what sarw.cpp:251-281 does serially. Chains are taken in increasing index; after a chain has
// been re-resolved, every already-resolved HIGHER chain near it is re-opened (its result assumed the old bead).
__device__ void cleanup_round(const GrowArgs& A, uint32_t round, int lane)
{
	volatile uint32_t* list = A.dirtyList;
	volatile uint32_t* count = &A.dirtyCount[round];
	const double R2 = independence_radius(A.g) * independence_radius(A.g);
	if (lane == 0) {   // chains put in conflict by the parallel part
		const uint32_t n2 = A.ctl->dirty2Count;
		for (uint32_t j = 0; j < n2; ++j) dirty_push_unique(list, count, list[A.g.nChains + j]);
		A.ctl->dirty2Count = 0;
	}
	__syncwarp();
	for (;;) {
		uint32_t chain = 0; int have = 0;
		if (lane == 0) have = dirty_pop_min(list, count, chain);
		have = __shfl_sync(FULL, have, 0); chain = __shfl_sync(FULL, chain, 0);
		if (!have) break;
		resolve_dirty_chain(A, round, chain, lane, [&](uint32_t c) { dirty_push_unique(list, count, c); });
		// re-open resolved higher chains whose neighbourhood just changed
		const uint32_t nd = *count;
		for (uint32_t j = lane; j < nd; j += 32) {
			const uint32_t e = list[j];
			if ((e & DONE_BIT) && (e & ~DONE_BIT) > chain && tip_dist2(A, e & ~DONE_BIT, chain) <= R2) list[j] = e & ~DONE_BIT;
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------ slab mode: exchange of one clean-up iteration (block 0 of every rank)
// message of one rank and iteration (clean_msg_bytes, sarw_host.h): {nRes, nPush, -, -}, nRes re-resolutions {x, y, z, trial, chain}, nPush chains put in conflict

// what resolve_dirty_chain did on the owner's side, repeated on a rank that only learns the result: the bead leaves / enters this window's
// cells (if its layer is held here), lpos, the chain's pending result and the replicated counters change exactly as on the owner
__device__ void slab_apply_remote_result(const GrowArgs& A, uint32_t round, const MrRecord& r)
{
	const Geom& g = A.g;
	const uint32_t T_ = (uint32_t)g.maxTrials, chain = r.pad, found = r.trial <= T_ ? r.trial : 0u;
	const uint32_t myId = round_id_base(g, round) + chain;
	const uint32_t cur = __ldcg(&A.resTrial[chain]), t0 = __ldcg(&A.origTrial[chain]);
	if (cur <= T_) remove_bead(g, A.T, myId);
	atomicAdd(&A.ctl->conflicts, 1ull);
	const long long before = cur <= T_ ? cur : T_, after = found ? found : T_;
	atomicAdd(&A.ctl->trials, (unsigned long long)(after - before));
	atomicAdd(&A.ctl->queries, (unsigned long long)(after - (long long)t0 + 1));
	const int hadBead = cur <= T_, hasBead = found != 0u, hadProg = cur < T_, hasProg = found != 0u && found < T_;
	if (hadBead != hasBead) atomicAdd(&A.roundAcc[round], (uint32_t)(hasBead - hadBead));
	if (hadProg != hasProg) atomicAdd(&A.roundProg[round], (uint32_t)(hasProg - hadProg));
	A.resTrial[chain] = found ? found : T_ + 1u;
	if (found) {
		const double p[3] = { r.x, r.y, r.z };
		for (int k = 0; k < 3; k++) A.resPos[3 * (size_t)chain + k] = p[k];
		insert_bead(g, A.T, p, myId);
	}
}

// Block 0: publish this rank's results and pushes to every rank, meet, apply the other ranks' results, update the conflict list.
// `scratch`: the block's dynamic shared memory (nothing else uses it during the clean-up).
__device__ void slab_clean_exchange(const GrowArgs& A, uint32_t round, uint32_t xseq, unsigned char* scratch)
{
	const uint32_t par = xseq & 1u, R = A.mrRanks, me = A.mrRank, cap = A.cleanCap;
	const size_t msgBytes = clean_msg_bytes(cap);
	const uint32_t nResRaw = __ldcg(&A.ctl->cleanResCount), nPushRaw = __ldcg(&A.ctl->cleanPushCount);
	const uint32_t nRes = min(nResRaw, cap), nPush = min(nPushRaw, cap);
	if (threadIdx.x == 0 && (nResRaw > cap || nPushRaw > cap)) atomicExch(&A.ctl->abort, 3u);   // more conflicts in one iteration than the message holds
	for (uint32_t p = 0; p < R; ++p) {
		unsigned char* msg = A.peerClean[p] + ((size_t)par * R + me) * msgBytes;
		uint4* dst = reinterpret_cast<uint4*>(msg + 16);
		const uint4* src = reinterpret_cast<const uint4*>(A.cleanRes);
		for (uint32_t j = threadIdx.x; j < 2u * nRes; j += blockDim.x) dst[j] = __ldcg(src + j);
		uint32_t* pd = reinterpret_cast<uint32_t*>(msg + 16 + (size_t)cap * sizeof(MrRecord));
		for (uint32_t j = threadIdx.x; j < nPush; j += blockDim.x) pd[j] = __ldcg(A.cleanPush + j);
		if (threadIdx.x == 0) *reinterpret_cast<uint4*>(msg) = make_uint4(nRes, nPush, 0u, 0u);
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		if (!mr_peer_meet_tag(A, par, xseq + 1u, 2u)) atomicExch(&A.ctl->abort, 1u);
		A.ctl->cleanResCount = 0u; A.ctl->cleanPushCount = 0u;
	}
	__syncthreads();
	const unsigned char* inbox = A.peerClean[me] + (size_t)par * R * msgBytes;
	// the other ranks' results: one thread each
	for (uint32_t p = 0; p < R; ++p) {
		if (p == me) continue;
		const unsigned char* msg = inbox + (size_t)p * msgBytes;
		const uint32_t n = min(__ldcg(reinterpret_cast<const uint32_t*>(msg)), cap);
		const uint4* rs = reinterpret_cast<const uint4*>(msg + 16);
		for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
			const uint4 a = __ldcg(rs + 2 * j), b = __ldcg(rs + 2 * j + 1);
			MrRecord r;
			r.x = __hiloint2double((int)a.y, (int)a.x); r.y = __hiloint2double((int)a.w, (int)a.z); r.z = __hiloint2double((int)b.y, (int)b.x); r.trial = b.z; r.pad = b.w;
			slab_apply_remote_result(A, round, r);
		}
	}
	__syncthreads();
	// list update, identical on every rank. Pass 1 (parallel over the list): close what was just resolved, re-open closed entries of
	// higher chains near a re-resolved chain. The re-resolved chains of all ranks are staged in shared memory in chunks.
	struct ResTip { double x, y, z; uint32_t chain, pad; };
	ResTip* tips = reinterpret_cast<ResTip*>(scratch);
	constexpr uint32_t CHUNK = 1024;
	uint32_t* openCount = reinterpret_cast<uint32_t*>(tips + CHUNK);
	if (threadIdx.x == 0) *openCount = 0u;
	__syncthreads();
	const uint32_t total = __ldcg(&A.dirtyCount[round]);
	const double R2 = independence_radius(A.g) * independence_radius(A.g);
	uint32_t allRes = 0;
	for (uint32_t p = 0; p < R; ++p) allRes += min(__ldcg(reinterpret_cast<const uint32_t*>(inbox + (size_t)p * msgBytes)), cap);
	// per-thread state of the entries this thread looks after (entry e = threadIdx.x + k * blockDim.x): bit k of `closeMask` / `openMask`
	for (uint32_t e0 = 0; e0 < total; e0 += blockDim.x) {
		const uint32_t e = e0 + threadIdx.x;
		uint32_t ent = e < total ? __ldcg(&A.dirtyList[e]) : 0u;
		const uint32_t chain = ent & ~DONE_BIT;
		double tip[3] = { 0, 0, 0 };
		if (e < total) for (int k = 0; k < 3; k++) tip[k] = __ldcg(&A.chains[chain].last[k]);
		bool resolvedNow = false, reopen = false;
		for (uint32_t c0 = 0; c0 < allRes; c0 += CHUNK) {
			__syncthreads();
			{	// stage re-resolved chains [c0, c0 + CHUNK) of the concatenated messages
				uint32_t skip = c0, off = 0;
				for (uint32_t p = 0; p < R; ++p) {
					const unsigned char* msg = inbox + (size_t)p * msgBytes;
					const uint32_t n = min(__ldcg(reinterpret_cast<const uint32_t*>(msg)), cap);
					if (skip >= n) { skip -= n; continue; }
					const uint4* rs = reinterpret_cast<const uint4*>(msg + 16);
					for (uint32_t j = skip + threadIdx.x; j < n && off + (j - skip) < CHUNK; j += blockDim.x) {
						const uint32_t ch = __ldcg(rs + 2 * j + 1).w;
						ResTip t; t.chain = ch; t.pad = 0;
						t.x = __ldcg(&A.chains[ch].last[0]); t.y = __ldcg(&A.chains[ch].last[1]); t.z = __ldcg(&A.chains[ch].last[2]);
						tips[off + (j - skip)] = t;
					}
					off += n - skip; skip = 0;
					if (off >= CHUNK) break;
				}
			}
			__syncthreads();
			const uint32_t m = min(CHUNK, allRes - c0);
			if (e < total)
				for (uint32_t j = 0; j < m; ++j) {
					const ResTip t = tips[j];
					if (t.chain == chain) { resolvedNow = true; continue; }
					if (t.chain < chain) {
						double d2 = 0;
						const double dd[3] = { tip[0] - t.x, tip[1] - t.y, tip[2] - t.z };
						for (int k = 0; k < 3; k++) { double d = dd[k]; d -= A.g.L[k] * rint(d * A.g.Linv[k]); d2 += d * d; }
						if (d2 <= R2) reopen = true;
					}
				}
		}
		if (e < total) {
			const bool done = ((ent & DONE_BIT) != 0u || resolvedNow) && !reopen;
			A.dirtyList[e] = chain | (done ? DONE_BIT : 0u);
			if (!done) atomicAdd(openCount, 1u);
		}
	}
	__syncthreads();
	// pass 2 (one thread): chains put in conflict by the re-resolutions, in rank order; a listed chain is opened again, a new one appended
	if (threadIdx.x == 0) {
		uint32_t open = *openCount, n = total;
		for (uint32_t p = 0; p < R; ++p) {
			const unsigned char* msg = inbox + (size_t)p * msgBytes;
			const uint32_t np = min(__ldcg(reinterpret_cast<const uint32_t*>(msg) + 1), cap);
			const uint32_t* pd = reinterpret_cast<const uint32_t*>(msg + 16 + (size_t)cap * sizeof(MrRecord));
			for (uint32_t j = 0; j < np; ++j) {
				const uint32_t c = __ldcg(pd + j);
				uint32_t at = n;
				for (uint32_t k = 0; k < n; ++k) if ((A.dirtyList[k] & ~DONE_BIT) == c) { at = k; break; }
				if (at == n) { if (n < A.g.nChains) { A.dirtyList[n++] = c; open++; } }
				else if (A.dirtyList[at] & DONE_BIT) { A.dirtyList[at] = c; open++; }
			}
		}
		A.dirtyCount[round] = n;
		A.ctl->cleanOpen = open;
	}
	__syncthreads();
}

#include "sarw_cluster.cuh"

// barrier across every CTA taking part in a round: the hardware cluster barrier when the whole launch is ONE
// thread-block cluster (small chain counts: a few hundred cycles), else the software grid barrier (cooperative launch)
template <bool CLUSTER>
__device__ __forceinline__ void round_barrier(Control* ctl, unsigned& barTarget)
{
	if (CLUSTER) {
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	} else {
		grid_barrier(&ctl->barCount, barTarget, gridDim.x);
	}
}

// MR = false: single-GPU instantiation, every multi-rank path is compiled out (they cost the 128-register kernel 170 -> 510 bytes of
// spill traffic and 13 % of its speed when they were selected at run time); MR = true: the phases of a multi-rank round
template <int WARPS, bool MR>
__global__ void __launch_bounds__(WARPS * 32, SARW_GROW_MINBLOCKS) grow_kernel(const __grid_constant__ GrowArgs A)
{
	constexpr int GROW_WARPS = WARPS;
	extern __shared__ __align__(128) unsigned char growSmem[];
	// per warp: 8 KB of staged cells (four 4 x 4 x 4 boxes: the quad path stages the regions of four chains at once; the one-chain path
	// uses the first box and keeps its blocked arcs in the second), 4 KB of neighbour lists (4 x 64 or 1 x 256 beads), one mbarrier
	Cell* rawAll = reinterpret_cast<Cell*>(growSmem);
	float4* nbAll = reinterpret_cast<float4*>(rawAll + WARPS * GROW_RAW_CELLS);
	unsigned long long* mbarAll = reinterpret_cast<unsigned long long*>(nbAll + WARPS * GROW_NB_CAP);
	const Geom& g = A.g;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t gw = blockIdx.x * GROW_WARPS + warp, nW = gridDim.x * GROW_WARPS;
	Cell* raw = rawAll + warp * GROW_RAW_CELLS;
	float4* nb = nbAll + warp * GROW_NB_CAP;
	float4* arcs = reinterpret_cast<float4*>(raw + CL_REGION_CELLS);
	static_assert(GROW_ARC_CAP * sizeof(float4) <= CL_REGION_CELLS * sizeof(Cell), "the arcs alias the second staged box");
	void* mbar = mbarAll + warp;
	unsigned tmaParity = 0;
	if (lane == 0) mbar_init(mbar, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	const uint32_t T_ = (uint32_t)g.maxTrials;
	__shared__ unsigned barTarget;   // grid barrier: arrivals needed for the next barrier of this launch (the wrapper zeroes the counter); thread 0 only
	if (threadIdx.x == 0) barTarget = 0;

	uint32_t round = __ldcg(&A.ctl->round);
	unsigned long long nBeads = __ldcg(&A.ctl->nBeads);
	uint32_t lastProg = __ldcg(&A.ctl->lastProgressed);
	const bool go = __ldcg(&A.ctl->initiated) != 0u;   // main() only enters the loop when initiateChain succeeded (sarw.cpp:66); checked here so that seeding and growth can be enqueued back to back
	unsigned long long myTrials = 0, myQueries = 0;
	const uint32_t phase = MR ? A.mrPhase : 0u;
	bool pendingValid = (phase == 1u);   // is there a round whose results are not yet committed? (always possible between phase-1/2 launches)
	const bool mr = MR && phase != 0u;
	const bool slab = MR && phase == 5u;   // slab decomposition: ownership by the tip's cell layer (see GrowArgs)
	uint32_t xseq = slab ? __ldcg(&A.ctl->xseq) : 0u;   // clean-up exchanges made so far (the same on every rank)
	if (phase == 3u) {       // multi-rank: commit the last round, nothing else
		for (uint32_t i = gw; i < g.nChains; i += nW) { ChainState cs = load_chain(A.chains + i); commit_pending(A, i, round - 1, cs, lane); }
		return;
	}

	const bool timer = (blockIdx.x == 0 && threadIdx.x == 0);
	long long cycA = 0, cycV = 0, cycC = 0; uint32_t cleanRounds = 0;
	long long sub[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, ts0 = 0, ts1 = 0;
#define SUBT(k) do { ts1 = clock64(); sub[k] += ts1 - ts0; ts0 = ts1; } while (0)
#define TRACE(stage) do { if (MR && A.trace && timer) { A.trace[0] = (round << 8) | (unsigned)(stage); __threadfence_system(); } } while (0)
	while (go && round < A.roundEnd && (long long)nBeads < g.target && lastProg) {
		const long long tk0 = clock64();
		ts0 = tk0;
		uint32_t start, end;
		round_range(g, nBeads, start, end);
		const uint32_t idRound = round_id_base(g, round);

		// ---------------- phase A: commit last round, then propose + tentatively insert
		if (mr && !slab && phase != 2u && pendingValid)          // chains proposed by other ranks: commit only, one thread each
			for (uint32_t i = (gw * 32 + lane); i < g.nChains; i += nW * 32)
				if ((i % A.mrRanks) != A.mrRank) commit_pending_thread(A, i, round - 1);
		uint32_t nOwnSlab = 0;
		TRACE(1);
		if (slab) {
			// slab mode: every chain is committed here, one thread each (the state is replicated), and goes to the rank whose slab holds the
			// cell layer of its new tip - computed from replicated data, so every rank arrives at the same assignment without a message
			for (uint32_t base = gw * 32; base < g.nChains; base += nW * 32) {
				const uint32_t i = base + lane;
				bool mine = false;
				if (i < g.nChains) {
					if (pendingValid) commit_pending_thread(A, i, round - 1);
					const double z = *reinterpret_cast<volatile const double*>(&A.chains[i].last[2]);
					const double tz = z * g.Linv[2];
					int cz = (int)floor((tz - floor(tz)) * g.sOverL[2]);   // the layer bead_cell puts the tip bead in
					cz = cz >= g.S[2] ? g.S[2] - 1 : (cz < 0 ? 0 : cz);
					const int o = slab_owner(g, cz);
					A.ownerOf[i] = (uint8_t)o;
					mine = (uint32_t)o == A.mrRank;
				}
				const unsigned m = __ballot_sync(FULL, mine);
				if (m) {
					uint32_t at = 0;
					if (lane == 0) at = atomicAdd(&A.ctl->ownCount[round & 1u], (uint32_t)__popc(m));
					at = __shfl_sync(FULL, at, 0);
					if (mine) A.ownList[at + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
				}
			}
			pendingValid = false;
			round_barrier<false>(A.ctl, barTarget);
			nOwnSlab = __ldcg(&A.ctl->ownCount[round & 1u]);
		}
		TRACE(2);
		if (phase != 2u)
		{
#if defined(SARW_NO_QUAD) || defined(SARW_PROF)
		// (diagnostics builds: one chain per warp, static chain -> warp mapping; whole-box contexts only, no slab mode)
		const uint32_t iFirst = mr ? A.mrRank + gw * A.mrRanks : gw, iStride = mr ? nW * A.mrRanks : nW;
		uint32_t pf = iFirst < g.nChains ? chain_prefetch(A, iFirst, lane) : 0u;
		for (uint32_t i = iFirst; i < g.nChains; i += iStride) {
#ifdef SARW_PROF
			long long prof_tp_ = clock64();
#endif
			ChainState cs = chain_open(A, i, pf, pendingValid, round - 1, lane);
			if (i + iStride < g.nChains) pf = chain_prefetch(A, i + iStride, lane);   // in flight while this chain is worked on
			PROF_MARK(48);   // state opened, last round committed
			if (i < start || i >= end) continue;
			if (cs.flags & CHAIN_DEAD) {   // permanently blocked: the reference burns maxTrials here every round
				if (lane == 0) {
					A.resTrial[i] = T_ + 1u; A.origTrial[i] = T_ + 1u;
					if (mr) { MrRecord r; r.x = r.y = r.z = 0; r.trial = T_ + 1u; r.pad = 0; mr_publish(A, i, round, r); }
				}
				continue;
			}
			Frame f; RegionPlan rp;
			prepare_chain(g, cs, f, rp);
			PROF_MARK(49);   // frame + region plan
			const bool hard = (cs.flags & CHAIN_HARD) != 0u;
			double cand[3] = { 0, 0, 0 }, vAcc[3] = { 0, 0, 0 };
			const uint32_t flagsBefore = cs.flags;
			const uint32_t accepted = propose_chain(g, A.T, A.tmap, A.ctl, A.chains, i, round, idRound, cs, f, rp, hard, raw, nb, GROW_NB_CAP, arcs, GROW_ARC_CAP, mbar, tmaParity, lane,
				cand, vAcc, myQueries);
			PROF_MARK(50);   // propose_chain
			const bool hardNow = accepted == 0u || accepted > 32u;
			cs.flags = (cs.flags & ~CHAIN_HARD) | (hardNow ? CHAIN_HARD : 0u);
			if (lane == 0) {
				if (cs.flags != flagsBefore) A.chains[i].flags = cs.flags;
				if (accepted) {
					insert_bead(g, A.T, cand, idRound + i);
					for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)i + k2] = cand[k2];
				}
				A.resTrial[i] = accepted ? accepted : T_ + 1u;
				A.origTrial[i] = accepted ? accepted : T_ + 1u;
				if (mr) { MrRecord r; r.x = cand[0]; r.y = cand[1]; r.z = cand[2]; r.trial = accepted ? accepted : T_ + 1u; r.pad = 0; mr_publish(A, i, round, r); }
			}
			__syncwarp();
			PROF_MARK(51);   // tentative insert + result stores (lane 0)
		}
#else
		// FOUR chains per warp at a time, eight lanes (= the first eight trials) each. Per chain-round the work of the one-chain path is
		// mostly warp-uniform fp64 arithmetic (state, frame, region) that all 32 lanes repeat, and nearly every chain accepts one of its
		// first trials: with 8 lanes per chain the same instruction stream serves four chains. A chain that is hard, whose region crosses
		// the periodic boundary or does not fit the staging buffers, or that finds nothing among its first eight trials, goes through
		// the one-chain path (propose_chain, all trials from the first) right after its quad: same walk, bit for bit.
		{
		const int sub = lane & 7, grp = lane >> 3, gbase = lane & 24;
		const uint32_t unit = mr ? A.mrRanks : 1u, first0 = mr ? A.mrRank : 0u;
		const uint32_t nOwn = slab ? nOwnSlab : (g.nChains > first0 ? (g.nChains - first0 + unit - 1u) / unit : 0u);   // chains this rank proposes
		Cell* rawq = raw + grp * CL_REGION_CELLS;
		float4* nbq = nb + grp * QUAD_NB_CAP;
		auto own_chain = [&](uint32_t q) { const uint32_t k = 4u * q + (uint32_t)grp; return slab ? __ldcg(A.ownList + k) : first0 + k * unit; };
		// prefetch of a chain's state: three 32-bit words per lane (words 0-15 the chain state, 16-21 last round's position, 22 its trial)
		auto prefetch = [&](uint32_t q, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
			w0 = w1 = w2 = 0u;
			if (4u * q + (uint32_t)grp < nOwn) {
				const uint32_t i = own_chain(q);
				const uint32_t* st = reinterpret_cast<const uint32_t*>(A.chains + i);
				w0 = __ldcg(st + sub); w1 = __ldcg(st + 8 + sub);
				if (sub < 6) w2 = __ldcg(reinterpret_cast<const uint32_t*>(A.resPos + 3 * (size_t)i) + sub);
				else if (sub == 6) w2 = __ldcg(A.resTrial + i);
			}
		};
		uint32_t pw0, pw1, pw2;
		prefetch(gw, pw0, pw1, pw2);
#ifdef SARW_QUAD_STATS
		long long qt0 = clock64(), qt1, qsec[13] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
#define QMARK(k) do { qt1 = clock64(); qsec[k] += qt1 - qt0; qt0 = qt1; } while (0)
#else
#define QMARK(k) do { } while (0)
#endif
		for (uint32_t q = gw; 4u * q < nOwn; q += nW) {
			const bool valid = 4u * q + (uint32_t)grp < nOwn;
			const uint32_t i = valid ? own_chain(q) : 0u;
			// ---- open the chain: state from the prefetched words, last round's result committed (commit_pending's effect)
			ChainState cs;
#define GSH(v, s) __shfl_sync(FULL, (v), gbase | (s))
#define GDBL(v, s) __hiloint2double((int)GSH(v, (s) + 1), (int)GSH(v, (s)))
			cs.last[0] = GDBL(pw0, 0); cs.last[1] = GDBL(pw0, 2); cs.last[2] = GDBL(pw0, 4);
			cs.pen[0] = GDBL(pw0, 6); cs.pen[1] = GDBL(pw1, 0); cs.pen[2] = GDBL(pw1, 2);
			cs.lastId = GSH(pw1, 4); cs.penId = GSH(pw1, 5); cs.length = GSH(pw1, 6); cs.flags = GSH(pw1, 7);
			{
				const double r0 = GDBL(pw2, 0), r1 = GDBL(pw2, 2), r2 = GDBL(pw2, 4);
				const uint32_t t = GSH(pw2, 6);
				if (pendingValid && valid) {
					if (t >= 1u && t <= T_) {
						const uint32_t id = round_id_base(g, round - 1) + i;
						cs.pen[0] = cs.last[0]; cs.pen[1] = cs.last[1]; cs.pen[2] = cs.last[2];
						cs.last[0] = r0; cs.last[1] = r1; cs.last[2] = r2;
						if (sub == 0) A.prev[id - g.idAtomBase] = cs.lastId;
						cs.penId = cs.lastId; cs.lastId = id; cs.length += 1;
						if (sub == 0) A.chains[i] = cs;
					}
					if (t != 0u && sub == 0) A.resTrial[i] = 0;
				}
			}
			prefetch(q + nW, pw0, pw1, pw2);                           // in flight while this quad is worked on
			QMARK(0);
			const bool inRange = valid && i >= start && i < end;
			const bool dead = inRange && (cs.flags & CHAIN_DEAD) != 0u;
			if (dead && sub == 0) {   // permanently blocked: the reference burns maxTrials here every round
				A.resTrial[i] = T_ + 1u; A.origTrial[i] = T_ + 1u;
				if (mr) { MrRecord r; r.x = r.y = r.z = 0; r.trial = T_ + 1u; r.pad = 0; mr_publish(A, i, round, r); }
			}
			const bool active = inRange && !dead;
			Frame f; RegionPlan rp;
			prepare_chain(g, cs, f, rp);                                  // (harmless arithmetic on the zero state of an invalid group)
			QMARK(1);
			// ---- stage the four regions: one tensor copy per chain, one mbarrier phase per warp
			const bool stage = active && !(cs.flags & CHAIN_HARD) && rp.tile != 0;
			const unsigned stageMask = __ballot_sync(FULL, stage && sub == 0);
			if (stageMask) {
				if (lane == 0) mbar_expect_tx(mbar, (unsigned)__popc(stageMask) * (unsigned)(CL_REGION_CELLS * sizeof(Cell)));
				__syncwarp();
				if (stage && sub == 0) tma_load_box3d(rawq, A.tmap, rp.lo[0] * (int)(sizeof(Cell) / 4), rp.lo[1], rp.lo[2] - g.tmapZ0, mbar);
			}
			// ---- while the copies fly: the exact candidate of trial sub + 1 (reference arithmetic)
			double candL[3] = { 0, 0, 0 }, vL[3] = { 0, 0, 0 }; float rel[3] = { 0, 0, 0 }; bool okL = false;
			if (stage) {
				const uint4 w = philox4x32_10(make_uint4(i, round, (uint32_t)sub >> 2, 0u), g.key0, g.key1);
				cone_candidate(g, f, cs.last, (double)pick(w, (uint32_t)sub & 3u) * (1.0 / 4294967296.0), candL);
				okL = (uint32_t)sub < T_ && !growth_rules_reject(g, candL, cs.last);
				to_frac(g, candL, vL);
				rel[0] = (float)(candL[0] - cs.last[0]); rel[1] = (float)(candL[1] - cs.last[1]); rel[2] = (float)(candL[2] - cs.last[2]);
			}
			int tot = 0; bool fits = stage;
			QMARK(2);
			if (stageMask) {
				mbar_wait(mbar, tmaParity); tmaParity ^= 1u;
				QMARK(3);
				// ---- neighbour list of each chain: lane `sub` takes box cells sub, sub + 8, ... (ix = sub & 3, iy = (sub >> 2) + 2 (k & 1), iz = k >> 1)
				const int ix = sub & 3, iy0 = sub >> 2;
				const bool xin = stage && ix < rp.nn[0];
				// pass 1, branch-free: which slots of the lane's eight cells hold a bead that counts (bit 3k + s), which cells are flagged
				uint32_t vm = 0; unsigned flgMask = 0;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const int iy = iy0 + 2 * (k & 1), iz = k >> 1;
					const uint32_t* cw = reinterpret_cast<const uint32_t*>(rawq + (iz * REGION_BOX + iy) * REGION_BOX + ix);
					const uint4 lo4 = *reinterpret_cast<const uint4*>(cw);
					const uint32_t w7 = cw[7];
					const bool in = xin && iy < rp.nn[1] && iz < rp.nn[2];
					const uint32_t m = (lo4.x < idRound && lo4.x != cs.lastId ? 1u : 0u) | (lo4.y < idRound && lo4.y != cs.lastId ? 2u : 0u) | (lo4.z < idRound && lo4.z != cs.lastId ? 4u : 0u);
					vm |= (in ? m : 0u) << (3 * k);
					flgMask |= (in && (w7 >> 16) != 0xFFFFu ? 1u : 0u) << k;
				}
				const int cnt = __popc(vm);
				QMARK(9);
				int incl = cnt;
#pragma unroll
				for (int d = 1; d < 8; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d, 8); if (sub >= d) incl += v; }
				tot = __shfl_sync(FULL, incl, 7, 8);
				fits = stage && tot <= QUAD_NB_CAP;
				int at = incl - cnt;
				if (!fits) vm = 0;
				// pass 2: one bead per iteration (a lane holds 2.5 on average)
				while (vm) {
					const int bit = __ffs((int)vm) - 1;
					vm &= vm - 1u;
					const int k = (bit * 11) >> 5, s2 = bit - 3 * k;             // bit / 3 for bit < 24
					const int iy = iy0 + 2 * (k & 1), iz = k >> 1;
					const Cell* cell = rawq + (iz * REGION_BOX + iy) * REGION_BOX + ix;
					const uint32_t id = cell->id[s2];
					const unsigned short* o = &cell->off[s2][0];
					const float sx = rp.base[0] + (float)ix * (float)g.a[0], sy = rp.base[1] + (float)iy * (float)g.a[1], sz = rp.base[2] + (float)iz * (float)g.a[2];
					nbq[at++] = make_float4(sx + dec_off(g, o[0], 0), sy + dec_off(g, o[1], 1), sz + dec_off(g, o[2], 2), __uint_as_float(id));
				}
				QMARK(10);
				// beads of flagged cells live in the overflow table: one lane of a group at a time appends them (a region holds a flagged cell in ~10 % of the chain-rounds of a melt)
				unsigned fm = __ballot_sync(FULL, fits && flgMask != 0u);
				for (int s2 = 0; s2 < 8 && fm; ++s2) {
					const unsigned has = fm & (0x01010101u << s2);
					if (!has) continue;
					int add = 0;
					if (fits && flgMask != 0u && sub == s2) {
#pragma unroll 1
						for (int k = 0; k < 8; k++) {
							if (!((flgMask >> k) & 1u)) continue;
							const int iy = iy0 + 2 * (k & 1), iz = k >> 1;
							const float sx = rp.base[0] + (float)ix * (float)g.a[0], sy = rp.base[1] + (float)iy * (float)g.a[1], sz = rp.base[2] + (float)iz * (float)g.a[2];
							ovf_for_cell(A.T, cell_index(g, rp.lo[0] + ix, rp.lo[1] + iy, rp.lo[2] + iz), [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
								if (id < idRound && id != cs.lastId) {
									if (tot + add < QUAD_NB_CAP) nbq[tot + add] = make_float4(sx + dec_off(g, o0, 0), sy + dec_off(g, o1, 1), sz + dec_off(g, o2, 2), __uint_as_float(id));
									add++;
								}
								return false;
							});
						}
					}
					tot += __shfl_sync(FULL, add, s2, 8);
					fm &= ~has;
				}
				fits = stage && tot <= QUAD_NB_CAP;
				__syncwarp();
				QMARK(4);
				// ---- lane-per-trial test against the chain's list (fp32 pre-filter, exact arithmetic inside the undecided band)
				if (okL && fits) {
					float dmin = 3.0e38f; int inBand = 0;
#pragma unroll 4
					for (int j = 0; j < tot; ++j) {
						const float4 b = nbq[j];
						const float dx = rel[0] - b.x, dy = rel[1] - b.y, dz = rel[2] - b.z;
						const float d2 = dx * dx + dy * dy + dz * dz;
						inBand += (d2 <= g.hi2);
						dmin = fminf(dmin, d2);
					}
					if (dmin < g.lo2) okL = false;
					else if (inBand) {
						for (int j = 0; j < tot && okL; ++j) {
							const float4 b = nbq[j];
							const float dx = rel[0] - b.x, dy = rel[1] - b.y, dz = rel[2] - b.z;
							const float d2 = dx * dx + dy * dy + dz * dz;
							if (d2 <= g.hi2 && exact_pair(g, vL, A.T.lpos + 3 * (size_t)__float_as_uint(b.w))) okL = false;
						}
					}
				}
			}
			const unsigned passMask = __ballot_sync(FULL, okL && fits);
			QMARK(5);
			const unsigned gm = (passMask >> gbase) & 0xFFu;
			const uint32_t accepted = gm ? (uint32_t)__ffs((int)gm) : 0u;       // lowest passing trial of the chain (1-based)
			const int winner = gbase | (accepted ? (int)accepted - 1 : 0);
			double cand[3];
#pragma unroll
			for (int k = 0; k < 3; k++) cand[k] = __shfl_sync(FULL, candL[k], winner);
			if (lane == 0) myQueries += (unsigned long long)(8 * __popc(stageMask));
			if (accepted && sub == 0) {
				insert_bead(g, A.T, cand, idRound + i);
				for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)i + k2] = cand[k2];
				A.resTrial[i] = accepted; A.origTrial[i] = accepted;
				if (mr) { MrRecord r; r.x = cand[0]; r.y = cand[1]; r.z = cand[2]; r.trial = accepted; r.pad = 0; mr_publish(A, i, round, r); }
			}
			__syncwarp();
			QMARK(6);
			// ---- the chains that need the one-chain path, one after the other with the whole warp
			unsigned slowMask = __ballot_sync(FULL, active && !accepted && sub == 0);
#ifdef SARW_QUAD_STATS
			if (sub == 0 && active && blockIdx.x == 0) {   // diagnostics (block 0 only): why chains leave the quad path (read with sarw_debug_paths)
				const int why = accepted ? 0 : ((cs.flags & CHAIN_HARD) ? 1 : (!rp.tile ? 2 : (!fits ? 3 : 4)));
				atomicAdd(&A.ctl->pathCnt[why < 4 ? why : 3], 1ull);
				if (why == 4) atomicAdd(&A.ctl->pathCyc[0], 1ull);
			}
#endif
			while (slowMask) {
				const int src = __ffs((int)slowMask) - 1;
				slowMask &= slowMask - 1u;
				const uint32_t i1 = __shfl_sync(FULL, i, src);
				ChainState c1;
#pragma unroll
				for (int k = 0; k < 3; k++) { c1.last[k] = __shfl_sync(FULL, cs.last[k], src); c1.pen[k] = __shfl_sync(FULL, cs.pen[k], src); }
				c1.lastId = __shfl_sync(FULL, cs.lastId, src); c1.penId = __shfl_sync(FULL, cs.penId, src);
				c1.length = __shfl_sync(FULL, cs.length, src); c1.flags = __shfl_sync(FULL, cs.flags, src);
				Frame f1; RegionPlan rp1;
				prepare_chain(g, c1, f1, rp1);
				const bool hard = (c1.flags & CHAIN_HARD) != 0u;
				double cand1[3] = { 0, 0, 0 }, vAcc[3] = { 0, 0, 0 };
				const uint32_t flagsBefore = c1.flags;
				const uint32_t acc1 = propose_chain(g, A.T, A.tmap, A.ctl, A.chains, i1, round, idRound, c1, f1, rp1, hard, raw, nb, GROW_NB_CAP, arcs, GROW_ARC_CAP, mbar, tmaParity, lane,
					cand1, vAcc, myQueries);
				const bool hardNow = acc1 == 0u || acc1 > 32u;
				c1.flags = (c1.flags & ~CHAIN_HARD) | (hardNow ? CHAIN_HARD : 0u);
				if (lane == 0) {
					if (c1.flags != flagsBefore) A.chains[i1].flags = c1.flags;
					if (acc1) {
						insert_bead(g, A.T, cand1, idRound + i1);
						for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)i1 + k2] = cand1[k2];
					}
					A.resTrial[i1] = acc1 ? acc1 : T_ + 1u;
					A.origTrial[i1] = acc1 ? acc1 : T_ + 1u;
					if (mr) { MrRecord r; r.x = cand1[0]; r.y = cand1[1]; r.z = cand1[2]; r.trial = acc1 ? acc1 : T_ + 1u; r.pad = 0; mr_publish(A, i1, round, r); }
				}
				__syncwarp();
			}
			QMARK(7);
#ifdef SARW_QUAD_STATS
			qsec[8] += 1;
#endif
#undef GSH
#undef GDBL
		}
#ifdef SARW_QUAD_STATS
		if (timer) for (int k2 = 0; k2 < 13; k2++) A.ctl->busy[k2] += (unsigned long long)qsec[k2];
#endif
#undef QMARK
		}
#endif
		}
		SUBT(0);
		TRACE(3);
		if (phase == 1u) break;                                  // the host all-gathers the records, then launches phase 2
		const MrRecord* allRec = A.mrAll;
		if (phase == 4u || slab) {                               // exchange inside the kernel: peer stores + flag handshake
			round_barrier<false>(A.ctl, barTarget);                            // every record store of this GPU has been issued
			SUBT(1);
			if (gw == 0 && lane == 0 && !mr_peer_meet(A, round)) atomicExch(&A.ctl->abort, 1u);
			SUBT(2);
			round_barrier<false>(A.ctl, barTarget);
			SUBT(3);
			if (__ldcg(&A.ctl->abort)) break;
			allRec = A.peerRec[A.mrRank] + (slab ? (size_t)(round & 1u) * g.nChains : (size_t)(round & 1u) * A.mrRanks * A.mrK);
		}
		TRACE(4);
		if (slab) {                                              // tentative beads of the chains other ranks own: the state of every chain, the beads of this window
			if (gw == 0 && lane == 0) A.ctl->ownCount[(round + 1u) & 1u] = 0u;
			for (uint32_t i = start + gw * 32 + lane; i < end; i += nW * 32) {
				if ((uint32_t)A.ownerOf[i] == A.mrRank) continue;
				const MrRecord* rp_ = allRec + i;
				MrRecord r;
				r.x = __ldcg(&rp_->x); r.y = __ldcg(&rp_->y); r.z = __ldcg(&rp_->z); r.trial = __ldcg(&rp_->trial); r.pad = 0;
				A.resTrial[i] = r.trial; A.origTrial[i] = r.trial;
				if (r.trial >= 1u && r.trial <= T_) {
					const double p[3] = { r.x, r.y, r.z };
					insert_bead(g, A.T, p, idRound + i);     // windowed: lpos always, the cell only if its layer is held here
					for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)i + k2] = p[k2];
				}
			}
		}
		if (phase == 2u || phase == 4u)                     // tentative beads proposed by the other ranks
			for (uint32_t i = start + gw * 32 + lane; i < end; i += nW * 32) {
				if ((i % A.mrRanks) == A.mrRank) continue;
				const MrRecord* rp_ = allRec + (size_t)(i % A.mrRanks) * A.mrK + i / A.mrRanks;
				MrRecord r;
				r.x = __ldcg(&rp_->x); r.y = __ldcg(&rp_->y); r.z = __ldcg(&rp_->z); r.trial = __ldcg(&rp_->trial); r.pad = 0;
				A.resTrial[i] = r.trial; A.origTrial[i] = r.trial;
				if (r.trial >= 1u && r.trial <= T_) {
					const double p[3] = { r.x, r.y, r.z };
					insert_bead(g, A.T, p, idRound + i);
					for (int k2 = 0; k2 < 3; k2++) A.resPos[3 * (size_t)i + k2] = p[k2];
				}
			}
		SUBT(4);
		pendingValid = true;
		round_barrier<false>(A.ctl, barTarget);
		SUBT(5);
		const long long tk1 = clock64();

		// ---------------- phase V: conflicts with tentative beads of lower-indexed chains of this round
		TRACE(5);
		const bool splitV = (phase == 4u && A.mrSplitVerify != 0u) || slab;
		{
			uint32_t acc = 0, prog = 0; unsigned long long tr = 0;
			if (splitV)   // round totals from the trial numbers every rank knows; the probing below covers this rank's chains only
				for (uint32_t i = start + gw * 32 + lane; i < end; i += nW * 32) {
					const uint32_t t = __ldcg(&A.resTrial[i]);
					if (t <= T_) { acc++; tr += t; if (t < T_) prog++; } else tr += T_;
				}
			const uint32_t stride = splitV ? A.mrRanks : 1u;
			const uint32_t first = splitV ? A.mrRank + ((start > A.mrRank ? start - A.mrRank + stride - 1u : 0u) / stride) * stride : start;   // first chain >= start of this rank
			// eight lanes per chain (one cell record each), four chains per warp and iteration, in the SAME chain -> warp mapping as the quads of
			// phase A: the pages of the cell table a warp touches here are the ones its SM touched a moment ago. (One THREAD per chain with its
			// eight records in flight at once - find_batch_kernel's scheme - measured 2.3x slower here: 32 chains per warp are ~80 distinct
			// 2 MB pages of the table, and the address-translation misses of one warp are served one after the other.)
			const int vsub = lane & 7, grp = lane >> 3;
			// (slab mode: the chains of this rank's list, the same list positions per warp as in phase A)
			const uint32_t vBegin = slab ? gw * 4u : first + gw * 4 * stride, vEnd = slab ? nOwnSlab : end, vStep = slab ? nW * 4u : nW * 4 * stride;
			for (uint32_t base = vBegin; base < vEnd; base += vStep) {
				uint32_t i = base + grp * stride;
				if (slab) { i = base + (uint32_t)grp < nOwnSlab ? __ldcg(A.ownList + base + grp) : end; if (i < start) i = end; }
				bool h = false; uint32_t t = 0;
				if (i < end) {
					t = __ldcg(&A.resTrial[i]);
					// position loaded together with the trial (always in bounds; garbage when there is no bead, then unused): one round trip, not two
					double pos[3] = { __ldcg(&A.resPos[3 * (size_t)i]), __ldcg(&A.resPos[3 * (size_t)i + 1]), __ldcg(&A.resPos[3 * (size_t)i + 2]) };
					if (t <= T_ && i > start) {
						double v[3];
						to_frac(g, pos, v);
						h = probe_cells<8>(g, A.T, v, idRound, idRound + i, ID_NONE, vsub);
					}
				}
				const unsigned bal = __ballot_sync(FULL, h);
				if (vsub == 0 && i < end) {
					if (!splitV) { if (t <= T_) { acc++; tr += t; if (t < T_) prog++; } else tr += T_; }
					if ((bal >> (grp * 8)) & 0xFFu) {
						uint32_t at = atomicAdd(&A.dirtyCount[round], 1u);
						A.dirtyList[at] = i;
					}
				}
			}
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				acc += __shfl_xor_sync(FULL, acc, d); prog += __shfl_xor_sync(FULL, prog, d); tr += __shfl_xor_sync(FULL, tr, d);
			}
			if (lane == 0) {
				if (acc) atomicAdd(&A.roundAcc[round], acc);
				if (prog) atomicAdd(&A.roundProg[round], prog);
				myTrials += tr;
			}
			SUBT(6);   // verify work of block 0 / warp 0 (the barrier that follows is in cycVerify)
		}
		TRACE(6);
		if (splitV) {   // exchange the conflict lists: block 0 stores this rank's list into every rank's inbox, ranks meet, lists are merged
			round_barrier<false>(A.ctl, barTarget);
			if (blockIdx.x == 0) {
				const uint32_t par = round & 1u, mine = min(__ldcg(&A.dirtyCount[round]), A.mrK);
				if (threadIdx.x == 0 && __ldcg(&A.dirtyCount[round]) > A.mrK) atomicExch(&A.ctl->abort, 2u);   // more conflicts than the inbox holds (cannot happen short of every chain conflicting)
				const size_t slot = ((size_t)par * A.mrRanks + A.mrRank) * (1u + A.mrK);
				for (uint32_t p = 0; p < A.mrRanks; ++p) {
					for (uint32_t j = threadIdx.x; j < mine; j += blockDim.x) A.peerDirty[p][slot + 1u + j] = __ldcg(&A.dirtyList[j]);
					if (threadIdx.x == 0) A.peerDirty[p][slot] = mine;
				}
				__threadfence_system();
				__syncthreads();
				if (threadIdx.x == 0 && !mr_peer_meet(A, round, 1u)) atomicExch(&A.ctl->abort, 1u);
				__syncthreads();
				// merge (rank order; the clean-up does not depend on the order of the list)
				uint32_t total = 0;
				for (uint32_t p = 0; p < A.mrRanks; ++p) {
					const uint32_t* in = A.peerDirty[A.mrRank] + ((size_t)par * A.mrRanks + p) * (1u + A.mrK);
					const uint32_t cnt = __ldcg(in);
					for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) A.dirtyList[total + j] = __ldcg(in + 1u + j);
					total += cnt;
				}
				__syncthreads();
				if (threadIdx.x == 0) A.dirtyCount[round] = total;
			}
		}
		round_barrier<false>(A.ctl, barTarget);
		if (splitV && __ldcg(&A.ctl->abort)) break;

		const long long tk2 = clock64();
		TRACE(7);
		// ---------------- clean-up (rare): serial re-resolution in chain order
		if (slab) {
			// slab mode: no rank holds all the cells, so conflicts are settled by fixed-point iteration instead of one ordered pass. Every open
			// entry of the (replicated) conflict list is re-resolved by the rank that owns the chain, all at once; the results and the chains
			// they put in conflict are exchanged; every rank applies them to its window and its replicated state and updates the list in the
			// same way: resolved entries are closed, every closed entry of a HIGHER chain near a re-resolved one is opened again (its result
			// assumed the old bead). Each iteration settles at least the lowest open chain for good, and the state it ends in - every chain
			// on the first trial that passes against the old beads and the final beads of all lower chains - is the serial loop's
			// (sarw.cpp:251-281) by induction over the chain index.
			uint32_t total = __ldcg(&A.dirtyCount[round]), open = total;
			while (open) {
				for (uint32_t e = gw; e < total; e += nW) {
					const uint32_t ent = __ldcg(&A.dirtyList[e]);
					if ((ent & DONE_BIT) || (uint32_t)A.ownerOf[ent] != A.mrRank) continue;
					resolve_dirty_chain(A, round, ent, lane, [&](uint32_t c) { const uint32_t at = atomicAdd(&A.ctl->cleanPushCount, 1u); if (at < A.cleanCap) A.cleanPush[at] = c; });
					if (lane == 0) {
						const uint32_t at = atomicAdd(&A.ctl->cleanResCount, 1u);
						if (at < A.cleanCap) {
							MrRecord r;
							r.x = A.resPos[3 * (size_t)ent]; r.y = A.resPos[3 * (size_t)ent + 1]; r.z = A.resPos[3 * (size_t)ent + 2]; r.trial = A.resTrial[ent]; r.pad = ent;
							A.cleanRes[at] = r;
						}
					}
					__syncwarp();
				}
				TRACE(8);
				round_barrier<false>(A.ctl, barTarget);
				TRACE(9);
				if (blockIdx.x == 0) slab_clean_exchange(A, round, xseq, growSmem);
				xseq++;
				round_barrier<false>(A.ctl, barTarget);
				if (__ldcg(&A.ctl->abort)) break;
				total = __ldcg(&A.dirtyCount[round]); open = __ldcg(&A.ctl->cleanOpen);
			}
			if (total) cleanRounds++;
			if (__ldcg(&A.ctl->abort)) break;
		} else if (__ldcg(&A.dirtyCount[round]) != 0u) {
			cleanup_parallel(A, round, gw, nW, lane);          // independent conflicts, one warp each
			round_barrier<false>(A.ctl, barTarget);
			// the ordered pass and its barrier only when the parallel part left something (a flag nobody writes after the barrier)
			if (gw == 0 && lane == 0) A.ctl->needOrdered[(round + 1u) & 1u] = 0u;
			if (__ldcg(&A.ctl->needOrdered[round & 1u]) != 0u) {
				if (gw == 0) cleanup_round(A, round, lane);      // the rest, in chain order
				round_barrier<false>(A.ctl, barTarget);
			}
			cleanRounds++;
		}
		TRACE(10);
		cycA += tk1 - tk0; cycV += tk2 - tk1; cycC += clock64() - tk2;
		nBeads += __ldcg(&A.roundAcc[round]);
		lastProg = __ldcg(&A.roundProg[round]) != 0u;
		round++;
	}

	// commit the last round of this launch (multi-rank: the next phase-1 launch, or phase 3, commits)
	if (pendingValid && (phase == 0u || phase == 4u || phase == 5u))
		for (uint32_t i = gw; i < g.nChains; i += nW) {
			ChainState cs = load_chain(A.chains + i);
			commit_pending(A, i, round - 1, cs, lane);
		}
	if (lane == 0) {
		if (myTrials) atomicAdd(&A.ctl->trials, myTrials);
		if (myQueries) atomicAdd(&A.ctl->queries, myQueries);
	}
	if (gw == 0 && lane == 0) {
		A.ctl->round = round; A.ctl->nBeads = nBeads; A.ctl->lastProgressed = lastProg ? 1u : 0u;
		if (slab) A.ctl->xseq = xseq;
	}
	if (timer) { A.ctl->cycPropose += cycA; A.ctl->cycVerify += cycV; A.ctl->cycCleanup += cycC; A.ctl->cleanupRounds += cleanRounds; }
	if (timer) for (int k2 = 0; k2 < 8; k2++) A.ctl->cycSub[k2] += sub[k2];
#undef SUBT
#undef TRACE
}

// ------------------------------------------------------------------ K7: compaction into acceptance order
__global__ void flag_slots_kernel(const double* __restrict__ lpos, uint32_t idAtomBase, uint64_t nSlots, uint32_t* __restrict__ flags)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nSlots) return;
	flags[s] = (__double_as_longlong(lpos[3 * (size_t)(idAtomBase + s)]) != -1ll) ? 1u : 0u;
}

__global__ void gather_atoms_kernel(Geom g, const double* __restrict__ lpos, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ scan,
	const uint32_t* __restrict__ prev, const ChainState* __restrict__ chains, uint64_t nSlots, uint32_t nSeeded, int terminated,
	double* __restrict__ xyz, int32_t* __restrict__ chainID, int32_t* __restrict__ type, int64_t* __restrict__ bonds)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nSlots || !flags[s]) return;
	const uint32_t gidx = scan[s];
	const uint32_t id = g.idAtomBase + (uint32_t)s;
	const uint64_t nSeedSlots = 2ull * g.nChains;
	const uint32_t chain = s < nSeedSlots ? (uint32_t)(s >> 1) : (uint32_t)((s - nSeedSlots) % g.nChains);
	const bool first = s < nSeedSlots && !(s & 1);
	if (xyz) for (int k = 0; k < 3; k++) xyz[3 * (size_t)gidx + k] = lpos[3 * (size_t)id + k];
	if (chainID) chainID[gidx] = (int32_t)chain + 1;
	if (type) type[gidx] = (first || (terminated && chains[chain].lastId == id)) ? 1 : 2;   // sarw.cpp:351-352,262,365
	if (bonds && !first) {
		const uint64_t b = s < nSeedSlots ? chain : (uint64_t)gidx - nSeeded;              // listBonds order, sarw.cpp:309,272
		bonds[2 * b] = (int64_t)scan[prev[s] - g.idAtomBase] + 1;
		bonds[2 * b + 1] = (int64_t)gidx + 1;
	}
}

// ------------------------------------------------------------------ size-independent checks of large walks
// Order-independent checksum of the whole walk in slot space (slot = 2*chain + {0,1} for seeds, 2n + round*n + chain for grown
// beads: a function of the walk alone, so one GPU, several GPUs and the CPU oracle must all give the same four words):
//   out[0] = sum, out[1] = xor of h(slot, x, y, z bits) over accepted slots; out[2] = accepted slots; out[3] = sum of h2(h, slot of the bonded predecessor)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31; return z;
}
__global__ void __launch_bounds__(256) walk_checksum_kernel(const double* __restrict__ lpos, const uint32_t* __restrict__ prev, uint32_t idAtomBase, uint64_t nSlots,
	unsigned long long* __restrict__ out)
{
	unsigned long long sum = 0, x = 0, cnt = 0, bsum = 0;
	for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nSlots; s += (uint64_t)gridDim.x * blockDim.x) {
		const double* p = lpos + 3 * (size_t)(idAtomBase + s);
		const unsigned long long bx = (unsigned long long)__double_as_longlong(p[0]);
		if (bx == ~0ull) continue;   // never accepted
		unsigned long long h = mix64(s + 0x9E3779B97F4A7C15ull);
		h = mix64(h ^ bx); h = mix64(h ^ (unsigned long long)__double_as_longlong(p[1])); h = mix64(h ^ (unsigned long long)__double_as_longlong(p[2]));
		const uint32_t pv = prev[s];
		sum += h; x ^= h; cnt += 1;
		bsum += mix64(h ^ (pv == ID_NONE ? ~0ull : (unsigned long long)(pv - idAtomBase)));
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		sum += __shfl_xor_sync(FULL, sum, d); x ^= __shfl_xor_sync(FULL, x, d); cnt += __shfl_xor_sync(FULL, cnt, d); bsum += __shfl_xor_sync(FULL, bsum, d);
	}
	if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], sum); atomicXor(&out[1], x); atomicAdd(&out[2], cnt); atomicAdd(&out[3], bsum); }
}

// beads whose wrapped position lies inside the axis-aligned box [lo, hi) (0 <= lo < hi <= L): position, lookup index, lookup index of the
// bonded predecessor (-1 for a chain's first bead) and chain; order is arbitrary. *count is the number found (may exceed cap: then truncated).
__global__ void __launch_bounds__(256) gather_box_kernel(Geom g, const double* __restrict__ lpos, const uint32_t* __restrict__ prev, uint64_t nSlots,
	double lo0, double lo1, double lo2, double hi0, double hi1, double hi2, unsigned long long cap, unsigned long long* __restrict__ count,
	double* __restrict__ xyz, int64_t* __restrict__ ids, int64_t* __restrict__ prevIds, int32_t* __restrict__ chainID)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nSlots) return;
	const double* p = lpos + 3 * (size_t)(g.idAtomBase + s);
	const double x = p[0], y = p[1], z = p[2];
	if (__double_as_longlong(x) == -1ll) return;
	const double wx = x - floor(x * g.Linv[0]) * g.L[0], wy = y - floor(y * g.Linv[1]) * g.L[1], wz = z - floor(z * g.Linv[2]) * g.L[2];
	if (wx < lo0 || wx >= hi0 || wy < lo1 || wy >= hi1 || wz < lo2 || wz >= hi2) return;
	const unsigned long long at = atomicAdd(count, 1ull);
	if (at >= cap) return;
	xyz[3 * at] = wx; xyz[3 * at + 1] = wy; xyz[3 * at + 2] = wz;
	ids[at] = (int64_t)s;
	const uint32_t pv = prev[s];
	prevIds[at] = pv == ID_NONE ? -1 : (int64_t)(pv - g.idAtomBase);
	const uint64_t nSeedSlots = 2ull * g.nChains;
	chainID[at] = (int32_t)(s < nSeedSlots ? (s >> 1) : ((s - nSeedSlots) % g.nChains)) + 1;
}

__global__ void chain_lengths_kernel(const ChainState* __restrict__ chains, uint32_t n, int32_t* __restrict__ out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (int32_t)chains[i].length;
}

__global__ void fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------ launch wrappers (called from sarw_abi.cpp)
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// `scratch`: find_scratch_bytes(g, n) bytes of device scratch for the layer-ordered processing of large batches, or nullptr
constexpr unsigned FIND_SORT_BLOCKS = 592;   // blocks of the two sort passes (4 per SM on B200); any number works
size_t find_scratch_bytes(const Geom& g, int64_t n)
{	return ((size_t)n * sizeof(SortedCand)) + (((size_t)n + (size_t)g.S[2] * (1 + FIND_SORT_BLOCKS)) * sizeof(uint32_t));
}
cudaError_t launch_find_batch(const Geom& g, const Tables& T, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit, int numSMs,
	void* scratch, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	const int64_t blocks = (n + 255) / 256;
	if (blocks > 0x7FFFFFFFll) return cudaErrorInvalidValue;
	const SortedCand* rec = nullptr; const uint32_t* perm = nullptr;
	if (scratch) {
		SortedCand* r = static_cast<SortedCand*>(scratch);
		uint32_t* p = reinterpret_cast<uint32_t*>(r + n);
		uint32_t* layerBase = p + n;
		const unsigned cb = FIND_SORT_BLOCKS;
		uint32_t* blockHist = layerBase + g.S[2];
		const int64_t chunk = (n + cb - 1) / cb;
		const size_t sm = (size_t)g.S[2] * sizeof(uint32_t);
		zbin_count_kernel<<<cb, 512, sm, st>>>(g, xyz, n, chunk, blockHist);
		zbin_rows_kernel<<<(unsigned)((g.S[2] + 7) / 8), 256, 0, st>>>(blockHist, layerBase, g.S[2], (int)cb);
		zbin_scan_kernel<<<1, 1024, 0, st>>>(layerBase, g.S[2]);
		zbin_scatter_kernel<<<cb, 512, sm, st>>>(g, xyz, ignore, n, chunk, blockHist, layerBase, r, p);
		rec = r; perm = p;
	}
	find_batch_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, T, xyz, ignore, n, hit, rec, perm);
	return cudaGetLastError();
}

cudaError_t launch_add_points(const Geom& g, const Tables& T, const double* xyz, int64_t n, uint32_t baseId, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	add_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, T, xyz, n, baseId);
	return cudaGetLastError();
}

cudaError_t launch_seed(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* prev, const double* grafts,
	Control* ctl, uint32_t* dirtyList, int numSMs, cudaStream_t st, int* launches)
{
	SeedArgs A{ g, T, chains, resTrial, prev, grafts, ctl, dirtyList };
	unsigned wblocks = (unsigned)((g.nChains + 7) / 8);
	unsigned cap = (unsigned)numSMs * 8;
	if (wblocks > cap) wblocks = cap;
	if (wblocks == 0) wblocks = 1;
	seed_propose_kernel<<<wblocks, 256, 0, st>>>(A);
	unsigned vblocks = (unsigned)((g.nChains + 31) / 32);
	if (vblocks > cap) vblocks = cap;
	if (vblocks == 0) vblocks = 1;
	seed_verify_kernel<<<vblocks, 256, 0, st>>>(A);
	seed_cleanup_kernel<<<1, 1, 0, st>>>(A);
	unsigned cblocks = (unsigned)((g.nChains + 255) / 256);
	if (cblocks > cap) cblocks = cap;
	if (cblocks == 0) cblocks = 1;
	seed_commit_kernel<<<cblocks, 256, 0, st>>>(A);
	*launches += 4;
	return cudaGetLastError();
}

constexpr size_t grow_smem_bytes(int warps)
{	return (size_t)warps * (GROW_RAW_CELLS * sizeof(Cell) + GROW_NB_CAP * sizeof(float4) + 8);
}

cudaError_t grow_max_blocks(int numSMs, int* maxBlocks)
{
	int perSM = 0;
	CK(cudaFuncSetAttribute(grow_kernel<GROW_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grow_smem_bytes(GROW_WARPS)));
	CK(cudaFuncSetAttribute(grow_kernel<GROW_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grow_smem_bytes(GROW_WARPS)));
	CK(cudaFuncSetAttribute(grow_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_smem_bytes(8)));
	CK(cudaFuncSetAttribute(grow_cluster_kernel<8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
	CK(cudaFuncSetAttribute(grow_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_smem_bytes(16)));
	CK(cudaFuncSetAttribute(grow_cluster_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, grow_kernel<GROW_WARPS, false>, GROW_WARPS * 32, grow_smem_bytes(GROW_WARPS)));
	*maxBlocks = perSM * numSMs;
	return cudaSuccess;
}

// largest chain count the single-cluster kernel covers (one chain per warp)
int grow_cluster_max_chains() { return CL_MAXCH; }

template <int WARPS>
static cudaError_t launch_cluster(const GrowArgs& A, cudaStream_t st)
{
	int ctas = (int)((A.g.nChains + WARPS - 1) / WARPS);
	if (ctas < 1) ctas = 1;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(WARPS * 32);
	cfg.dynamicSmemBytes = cluster_smem_bytes(WARPS); cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, grow_cluster_kernel<WARPS>, A);
}

#ifdef SARW_PROF
cudaError_t prof_read(unsigned long long* out, int n, bool reset)
{
	std::vector<unsigned long long> h((size_t)64 * PROF_ROWS);
	CK(cudaMemcpyFromSymbol(h.data(), g_prof, h.size() * sizeof(unsigned long long)));
	for (int k = 0; k < n && k < 64; k++) { unsigned long long s = 0; for (int r = 0; r < PROF_ROWS; r++) s += h[(size_t)64 * r + k]; out[k] = s; }
	if (reset) { std::fill(h.begin(), h.end(), 0ull); CK(cudaMemcpyToSymbol(g_prof, h.data(), h.size() * sizeof(unsigned long long))); }
	return cudaSuccess;
}
#else
cudaError_t prof_read(unsigned long long*, int, bool) { return cudaErrorNotSupported; }
#endif

// diagnostics: how many clusters of the size launch_cluster would use for nChains chains can be resident at once
int cluster_max_active(uint32_t nChains)
{
	const bool w8 = nChains <= 8u * CL_MAX_CTAS;
	const int W = w8 ? 8 : 16;
	int ctas = (int)((nChains + W - 1) / W); if (ctas < 1) ctas = 1;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(W * 32); cfg.dynamicSmemBytes = cluster_smem_bytes(W);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	int n = -1;
	cudaError_t e = w8 ? cudaOccupancyMaxActiveClusters(&n, grow_cluster_kernel<8>, &cfg) : cudaOccupancyMaxActiveClusters(&n, grow_cluster_kernel<16>, &cfg);
	return e == cudaSuccess ? n : -(int)e;
}

cudaError_t launch_grow(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* origTrial, double* resPos, uint32_t* prev, Control* ctl,
	uint32_t* roundAcc, uint32_t* roundProg, uint32_t* dirtyCount, uint32_t* dirtyList, uint32_t roundEnd, int blocks, int useCluster, cudaStream_t st)
{
	GrowArgs A{ g, T, chains, resTrial, origTrial, resPos, prev, ctl, roundAcc, roundProg, dirtyCount, dirtyList, roundEnd, 0u, 1u, 0u, 0u, nullptr, nullptr, {}, {}, {}, 0u,
		nullptr, nullptr, nullptr, nullptr, 0u, {}, nullptr, {} };
	if (g.useTmap && T.tmapHost) memcpy(A.tmap, T.tmapHost, sizeof A.tmap); else A.g.useTmap = 0;
	if (useCluster) {
		// the whole launch is one thread-block cluster: CTAs are co-scheduled by the hardware, synchronise with
		// barrier.cluster and exchange tentative beads through distributed shared memory. 8 warps per CTA (255
		// registers per thread) up to 128 chains, 16 warps per CTA beyond.
		return g.nChains <= 8u * CL_MAX_CTAS ? launch_cluster<8>(A, st) : launch_cluster<16>(A, st);
	}
	CK(cudaMemsetAsync(&ctl->barCount, 0, sizeof(unsigned), st));   // the grid barrier counts arrivals from zero in every launch
	void* args[] = { &A };
	return cudaLaunchCooperativeKernel((const void*)grow_kernel<GROW_WARPS, false>, dim3((unsigned)blocks), dim3(GROW_WARPS * 32), args,
		grow_smem_bytes(GROW_WARPS), st);
}

// one phase of a multi-rank round (grow_kernel with mrPhase != 0); always the cooperative grid kernel
unsigned* g_traceHost[8] = {};
cudaError_t launch_grow_mr(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* origTrial, double* resPos, uint32_t* prev, Control* ctl,
	uint32_t* roundAcc, uint32_t* roundProg, uint32_t* dirtyCount, uint32_t* dirtyList, uint32_t roundEnd, int blocks,
	uint32_t rank, uint32_t ranks, uint32_t phase, void* localRec, const void* allRec, void* const* peerRec, void* const* peerFlags,
	void* const* peerDirty, uint32_t splitVerify, const SlabBuffers* slab, cudaStream_t st)
{
	const uint32_t K = (g.nChains + ranks - 1) / ranks;
	GrowArgs A{ g, T, chains, resTrial, origTrial, resPos, prev, ctl, roundAcc, roundProg, dirtyCount, dirtyList, roundEnd, rank, ranks, phase, K,
		(MrRecord*)localRec, (const MrRecord*)allRec, {}, {}, {}, splitVerify, nullptr, nullptr, nullptr, nullptr, 0u, {}, nullptr, {} };
	if (phase == 5u) {
		if (!slab) return cudaErrorInvalidValue;
		A.ownList = slab->ownList; A.ownerOf = slab->ownerOf; A.cleanRes = (MrRecord*)slab->cleanRes; A.cleanPush = slab->cleanPush; A.cleanCap = slab->cleanCap;
		for (uint32_t p = 0; p < 8; ++p) A.peerClean[p] = p < ranks ? (unsigned char*)slab->peerClean[p] : nullptr;
	}
	A.g.winStrict = 1;   // growth: a probe outside this rank's window is a protocol error (counted in Control::oowCount)
	if (getenv("SARW_TRACE_MAPPED")) {   // diagnostics: where was block 0 when a launch died?
		static unsigned* hostTrace[8] = {};
		if (!hostTrace[rank & 7]) cudaHostAlloc((void**)&hostTrace[rank & 7], 256, cudaHostAllocMapped | cudaHostAllocPortable);
		unsigned* d = nullptr;
		if (hostTrace[rank & 7] && cudaHostGetDevicePointer((void**)&d, hostTrace[rank & 7], 0) == cudaSuccess) { A.trace = d; g_traceHost[rank & 7] = hostTrace[rank & 7]; }
	}
	if (g.useTmap && T.tmapHost) memcpy(A.tmap, T.tmapHost, sizeof A.tmap); else A.g.useTmap = 0;
	for (uint32_t p = 0; p < 8; ++p) {
		A.peerRec[p] = peerRec && p < ranks ? (MrRecord*)peerRec[p] : nullptr; A.peerFlags[p] = peerFlags && p < ranks ? (uint32_t*)peerFlags[p] : nullptr;
		A.peerDirty[p] = peerDirty && p < ranks ? (uint32_t*)peerDirty[p] : nullptr;
	}
	CK(cudaMemsetAsync(&ctl->barCount, 0, sizeof(unsigned), st));
	void* args[] = { &A };
	return cudaLaunchCooperativeKernel((const void*)grow_kernel<GROW_WARPS, true>, dim3((unsigned)blocks), dim3(GROW_WARPS * 32), args,
		grow_smem_bytes(GROW_WARPS), st);
}

cudaError_t launch_download(const Geom& g, const Tables& T, const uint32_t* prev, const ChainState* chains, uint64_t nSlots, uint32_t nSeeded,
	int terminated, uint32_t* flags, uint32_t* scan, void* cubTemp, size_t cubBytes,
	double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds, cudaStream_t st)
{
	if (nSlots == 0) return cudaSuccess;
	unsigned blocks = (unsigned)((nSlots + 255) / 256);
	flag_slots_kernel<<<blocks, 256, 0, st>>>(T.lpos, g.idAtomBase, nSlots, flags);
	CK(cub::DeviceScan::ExclusiveSum(cubTemp, cubBytes, flags, scan, (int64_t)nSlots, st));
	gather_atoms_kernel<<<blocks, 256, 0, st>>>(g, T.lpos, flags, scan, prev, chains, nSlots, nSeeded, terminated, xyz, chainID, type, bonds);
	return cudaGetLastError();
}

cudaError_t download_scan_bytes(uint64_t nSlots, size_t* bytes)
{
	*bytes = 0;
	return cub::DeviceScan::ExclusiveSum(nullptr, *bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)nSlots, 0);
}

cudaError_t launch_walk_checksum(const Tables& T, const uint32_t* prev, uint32_t idAtomBase, uint64_t nSlots, unsigned long long* out4, int numSMs, cudaStream_t st)
{
	CK(cudaMemsetAsync(out4, 0, 4 * sizeof(unsigned long long), st));
	if (!nSlots) return cudaSuccess;
	uint64_t blocks = (nSlots + 255) / 256; if (blocks > (uint64_t)numSMs * 16) blocks = (uint64_t)numSMs * 16;
	walk_checksum_kernel<<<(unsigned)blocks, 256, 0, st>>>(T.lpos, prev, idAtomBase, nSlots, out4);
	return cudaGetLastError();
}

cudaError_t launch_gather_box(const Geom& g, const Tables& T, const uint32_t* prev, uint64_t nSlots, const double lo[3], const double hi[3], uint64_t cap,
	unsigned long long* count, double* xyz, int64_t* ids, int64_t* prevIds, int32_t* chainID, cudaStream_t st)
{
	CK(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
	if (!nSlots) return cudaSuccess;
	const uint64_t blocks = (nSlots + 255) / 256;
	if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
	gather_box_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, T.lpos, prev, nSlots, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], cap, count, xyz, ids, prevIds, chainID);
	return cudaGetLastError();
}

cudaError_t launch_chain_lengths(const ChainState* chains, uint32_t n, int32_t* out, cudaStream_t st)
{
	if (!n) return cudaSuccess;
	chain_lengths_kernel<<<(n + 255) / 256, 256, 0, st>>>(chains, n, out);
	return cudaGetLastError();
}

cudaError_t launch_fill_u32(uint32_t* p, uint64_t n, uint32_t v, cudaStream_t st)
{
	if (!n) return cudaSuccess;
	uint64_t blocks = (n + 255) / 256; if (blocks > 65535) blocks = 65535;
	fill_u32_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, v);
	return cudaGetLastError();
}

} // namespace sarw
