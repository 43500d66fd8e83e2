// sarw_cluster.cuh — growth kernel for small chain counts (nChains <= 256): ONE thread-block cluster, one chain per
// warp for the whole run. Included by sarw_kernels.cu inside namespace sarw (it uses stage_region, eval_trial_coop,
// cleanup_round, load_chain ... defined there).
//
// Same round protocol and same results as grow_kernel (sarw.cpp:220-283 semantics), but a round touches global memory
// only to stage the neighbourhood (one dependent load round trip) and to store the accepted bead:
//   * chain state lives in registers of the owning warp for the whole launch;
//   * every warp broadcasts its tentative bead {xyz, trial, cell} into the shared memory of EVERY CTA of the cluster
//     (distributed shared memory), so after the hardware cluster barrier each warp checks conflicts with lower chains and
//     counts the round's accepted beads from its own CTA's copy, without global loads or atomics;
//   * beads enter the cell list AFTER that barrier with plain stores: the slot is the rank-th free slot of the cell as seen while
//     staging, rank = position among this round's beads of the same cell, which every warp derives from the same table — no
//     atomic round trip;
//   * the second cluster barrier publishes the stores and the (rare) conflict flag; conflicts fall back to the shared
//     serial clean-up (cleanup_round) through global memory.
#pragma once
// (sarw_kernels.cu includes <cooperative_groups.h> at file scope before opening namespace sarw)
namespace cgx = ::cooperative_groups;

constexpr int CL_MAX_CTAS = 16;         // non-portable cluster size
constexpr int CL_MAXCH = 16 * CL_MAX_CTAS;   // table capacity: 16 warps per CTA x 16 CTAs (the 8-warp variant uses half of it)
constexpr int CL_REGION_CELLS = 64;     // staged region: at most 4 x 4 x 4 cells of 32 bytes (2 KB per chain)


// tentative bead of one chain in one round, as broadcast to every CTA
struct __align__(8) TentRec {
	double x, y, z;
	uint32_t trial;     // 0 = chain not in this round's range, 1..T accepted trial, T+1 none
	int32_t cx, cy, cz; // wrapped cell coordinates of the bead
};
static_assert(sizeof(TentRec) == 40, "TentRec is 40 bytes");

// staged beads / blocked arcs kept per chain: the 8-warp variant has shared memory to spare
__host__ __device__ constexpr int cl_nb_cap(int warps) { return warps <= 8 ? 256 : 128; }
__host__ __device__ constexpr int cl_arc_cap(int warps) { return warps <= 8 ? 128 : 64; }
constexpr size_t cluster_smem_bytes(int warps)
{	return (size_t)warps * CL_REGION_CELLS * sizeof(Cell) + (size_t)warps * cl_nb_cap(warps) * sizeof(float4) + (size_t)warps * cl_arc_cap(warps) * sizeof(float4)
		+ 2 * (size_t)CL_MAXCH * sizeof(TentRec) + (size_t)warps * 8 + 16;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Blocked arcs as (cos alpha, sin alpha, kappa): a trial angle theta is CERTAINLY rejected by the bead iff
// cos(alpha) cos(theta) + sin(alpha) sin(theta) >= kappa. kappa carries margins (g.arcShrink = 1e-5 A + the quantisation error of
// the staged positions in distance, 3e-5 in cos) that
// dominate the fp32 evaluation error of both sides; arcs only ever skip trials, the exact test decides everything else.
__device__ __forceinline__ int build_arcs_cs(const Geom& g, const Frame& f, const double last[3], const float4* nb, int nnb, float4* arcs, int arcCap, int lane)
{
	const double rho = g.dSinPhi, h = g.dCosPhi;
	const double ts = g.thresh - g.arcShrink;
	const float halfInvRho = (float)(0.5 / rho);
	int narcs = 0;
	for (int jb = 0; jb < nnb; jb += 32) {
		const int j = jb + lane;
		bool have = false; float4 arc = make_float4(0.f, 0.f, 2.f, 0.f);
		if (j < nnb && ts > 0) {
			const float4 b = nb[j];
			if (b.x < 1e29f) {
				const double wx = (double)b.x - h * f.Z[0], wy = (double)b.y - h * f.Z[1], wz = (double)b.z - h * f.Z[2];
				const double px = wx * f.X[0] + wy * f.X[1] + wz * f.X[2];
				const double py = wx * f.Y[0] + wy * f.Y[1] + wz * f.Y[2];
				const double pz = wx * f.Z[0] + wy * f.Z[1] + wz * f.Z[2];
				const double R2 = px * px + py * py;
				if (R2 > 1e-8) {
					const float rinv = rsqrtf((float)R2);
					const float kappa = (float)(R2 + pz * pz + rho * rho - ts * ts) * halfInvRho * rinv;
					if (kappa < 1.0f) { have = true; arc = make_float4((float)px * rinv, (float)py * rinv, kappa + 3e-5f * (1.0f + fabsf(kappa)), 0.f); }
				}
			}
		}
		const unsigned m = __ballot_sync(FULL, have);
		if (narcs + __popc(m) > arcCap) return -1;
		if (have) arcs[narcs + __popc(m & ((1u << lane) - 1u))] = arc;
		narcs += __popc(m);
	}
	{	// the rules of sarw.cpp:201-211 and the growthBias extension have the same form: a coordinate (or bias.step) of the
		// candidate is base + rho*R*cos(theta - beta), so each rule certainly rejects an arc of theta. Lanes 0..3: lower wall,
		// upper wall, growthBias, zCenter half-space.
		bool have = false; float4 arc = make_float4(0.f, 0.f, 2.f, 0.f);
		const int wk = g.boundary == 1 ? 2 : (g.boundary == 2 ? 1 : (g.boundary == 3 ? 0 : -1));
		double cx = 0, cy = 0, base = 0, limit = 0; int kind = 0;   // kind 1: reject below `limit`, 2: reject above `limit`
		if (lane < 2 && wk >= 0) { cx = f.X[wk]; cy = f.Y[wk]; base = last[wk] + h * f.Z[wk]; limit = lane == 0 ? 0.0 : g.L[wk]; kind = lane == 0 ? 1 : 2; }
		else if (lane == 2 && g.biased) {
			cx = g.bias[0] * f.X[0] + g.bias[1] * f.X[1] + g.bias[2] * f.X[2]; cy = g.bias[0] * f.Y[0] + g.bias[1] * f.Y[1] + g.bias[2] * f.Y[2];
			base = h * (g.bias[0] * f.Z[0] + g.bias[1] * f.Z[1] + g.bias[2] * f.Z[2]); limit = 0.0; kind = 1;
		} else if (lane == 3 && g.graftLoc == 2) {
			const double lastPos = last[2] - 0.5 * g.L[2];
			cx = f.X[2]; cy = f.Y[2]; base = last[2] + h * f.Z[2]; limit = 0.5 * g.L[2];
			kind = lastPos > 0 ? 1 : (lastPos < 0 ? 2 : 0);
		}
		if (kind) {
			const double scale = (lane == 2) ? fabs(g.bias[0]) + fabs(g.bias[1]) + fabs(g.bias[2]) : 1.0;
			const double delta = 1e-5 * (scale > 1.0 ? scale : 1.0);
			const double R2 = cx * cx + cy * cy;
			// certainly rejected iff value < limit - delta (kind 1) or value > limit + delta (kind 2), value = base + rho*(cx cos + cy sin)
			if (R2 > 1e-16) {
				const double R = sqrt(R2), sgn = kind == 1 ? -1.0 : 1.0;
				// kind 1: cos(theta-beta) < (limit-delta-base)/(rho R)  <=>  cos(theta-(beta+pi)) > (base+delta-limit)/(rho R)
				const double kappa = (kind == 1 ? (base + delta - limit) : (limit + delta - base)) / (rho * R);
				if (kappa < 1.0) { have = true; arc = make_float4((float)(sgn * cx / R), (float)(sgn * cy / R), (float)kappa + 3e-5f * (1.0f + fabsf((float)kappa)), 0.f); }
			} else if ((kind == 1 && base < limit - delta) || (kind == 2 && base > limit + delta)) {
				have = true; arc = make_float4(1.f, 0.f, -2.f, 0.f);   // the coordinate does not depend on theta and violates: everything is rejected
			}
		}
		const unsigned m = __ballot_sync(FULL, have);
		if (narcs + __popc(m) > arcCap) return -1;
		if (have) arcs[narcs + __popc(m & ((1u << lane) - 1u))] = arc;
		narcs += __popc(m);
	}
	__syncwarp();
	return narcs;
}

// Do the (shrunk) arcs, together with the never-sampled gap [359 deg, 360 deg) of sarw.cpp:177, cover the whole circle?
// Closed arcs cover the circle iff every arc's upper end lies strictly inside another arc. Centre and half-width of every
// arc are computed once (acos / atan2 in double, one arc per lane) into `scratch`, then compared pairwise. Rare path.
__device__ __forceinline__ bool arcs_cs_cover_circle(const float4* arcs, int narcs, double2* scratch, int lane)
{
	if (narcs <= 0) return false;
	const double PI = 3.14159265358979323846;
	const double gapLo = 359.0 * PI / 180 - 1e-9, gapHi = 2 * PI;
	bool full = false;
	for (int i = lane; i <= narcs; i += 32) {
		double c, w;
		if (i < narcs) {
			const double k = (double)arcs[i].z;
			c = atan2((double)arcs[i].y, (double)arcs[i].x);
			w = k <= -1.0 ? PI : (k >= 1.0 ? 0.0 : acos(k) - 1e-6);
		} else { c = 0.5 * (gapLo + gapHi); w = 0.5 * (gapHi - gapLo) - 1e-7; }
		scratch[i] = make_double2(c, w);
		full = full || w >= PI;
	}
	__syncwarp();
	if (__any_sync(FULL, full)) return true;
	bool allCovered = true;
	for (int ib = 0; ib <= narcs; ib += 32) {
		const int i = ib + lane;
		bool covered = true;
		if (i <= narcs) {
			const double2 me = scratch[i];
			const double e = me.x + me.y;
			covered = false;
			for (int j = 0; j <= narcs && !covered; ++j) {
				const double2 o = scratch[j];
				double d = e - o.x;
				d -= 2 * PI * rint(d * (0.5 / PI));
				covered = fabs(d) < o.y - 1e-6;
			}
		}
		allCovered = allCovered && __all_sync(FULL, covered);
	}
	return allCovered;
}

// ---------------------------------------------------------------------------------------------------------------
// TMA staging: the region (<= 4x4x4 cells, x-rows contiguous in memory) is copied row by row into the warp's shared
// memory with cp.async.bulk (one lane per row, completion on a per-warp mbarrier). The copies cost no registers and
// no issue slots while in flight, so the warp prepares its exact candidates meanwhile.
// tile != 0: the region lies inside the table without crossing the periodic boundary and spans <= 4 cells per dimension: it is staged
// by ONE 3-D tensor copy as a fixed 4 x 4 x 4 box (row stride 4, layer stride 16); otherwise compact (row stride nn0, layer nn0*nn1)
struct RegionPlan { int lo[3], nn[3], ncell; float base[3]; int tile; };
constexpr int REGION_BOX = 4;
__device__ __forceinline__ int region_index(const RegionPlan& rp, int ix, int iy, int iz)
{	return rp.tile ? (iz * REGION_BOX + iy) * REGION_BOX + ix : (iz * rp.nn[1] + iy) * rp.nn[0] + ix;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* mbar, unsigned count)
{	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* mbar, unsigned bytes)
{	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* mbar, unsigned parity)
{
	asm volatile(
		"{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
		:: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dstSmem, const void* srcGlobal, unsigned bytes, void* mbar)
{	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// one instruction for the whole region: 3-D tiled tensor copy (coordinates: x in 4-byte elements, 32 per cell; y, z in cells)
__device__ __forceinline__ void tma_load_box3d(void* dstSmem, const void* tmap, int c0, int c1, int c2, void* mbar)
{	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(smem_u32(dstSmem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ int wrap_near(int c, int S)   // c is normally within one period of [0,S): no integer division
{	c += (c < 0) ? S : 0;
	c -= (c >= S) ? S : 0;
	if ((unsigned)c >= (unsigned)S) { c %= S; c += (c < 0) ? S : 0; }   // tiny boxes: several periods away
	return c;
}

// frame, wrapped tip and region of the chain's next proposal; depends only on the chain state
__device__ __forceinline__ void prepare_chain(const Geom& g, const ChainState& cs, Frame& f, RegionPlan& rp)
{
	cone_frame(cs.last, cs.pen, f);
	double vl[3];
	to_frac(g, cs.last, vl);
	rp.ncell = 1;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const double lw = vl[k] * g.L[k];
		// cell-space bounds of the cone circle +- reach, in fp32 with a margin that dominates the rounding (conservative)
		const float cenc = (float)((lw + g.dCosPhi * f.Z[k]) * g.ainv[k]);
		const float w = 1.0f - (float)f.Z[k] * (float)f.Z[k];
		const float ec = ((float)g.dSinPhi * sqrtf(w > 0.f ? w : 0.f) + (float)g.regionReach) * (float)g.ainv[k] + 1e-3f + 4e-7f * fabsf(cenc);
		rp.lo[k] = (int)floorf(cenc - ec);
		rp.nn[k] = (int)floorf(cenc + ec) - rp.lo[k] + 1;
		rp.ncell *= rp.nn[k];
		rp.base[k] = (float)((double)rp.lo[k] * g.a[k] - lw);
	}
	rp.tile = g.useTmap && rp.nn[0] <= REGION_BOX && rp.nn[1] <= REGION_BOX && rp.nn[2] <= REGION_BOX
		&& rp.lo[0] >= 0 && rp.lo[1] >= 0 && rp.lo[2] >= g.tmapZ0 && rp.lo[0] + rp.nn[0] <= g.S[0] && rp.lo[1] + rp.nn[1] <= g.S[1] && rp.lo[2] + rp.nn[2] <= g.tmapZ0 + g.tmapZn;
}

// one lane per x-row: bulk-copy its nn0 cells (split in two when the row crosses the periodic boundary)
__device__ __forceinline__ void region_issue_tma(const Geom& g, const Tables& T, const void* tmap, const RegionPlan& rp, Cell* raw, void* mbar, int lane)
{
	// (the generic-proxy reads of `raw` from the previous round completed long ago: their results were consumed and two
	// cluster barriers lie in between, so the async-proxy writes issued here cannot overtake them)
	if (rp.tile) {   // the usual case: one tensor copy of the 4 x 4 x 4 box (cells beyond the table edge arrive as zeros and are never read)
		if (lane == 0) {
			mbar_expect_tx(mbar, (unsigned)(REGION_BOX * REGION_BOX * REGION_BOX) * (unsigned)sizeof(Cell));
			tma_load_box3d(raw, tmap, rp.lo[0] * (int)(sizeof(Cell) / 4), rp.lo[1], rp.lo[2] - g.tmapZ0, mbar);
		}
		__syncwarp();
		return;
	}
	if (lane == 0) mbar_expect_tx(mbar, (unsigned)rp.ncell * (unsigned)sizeof(Cell));
	__syncwarp();
	const int rows = rp.nn[1] * rp.nn[2];
	if (lane < rows) {
		const int iz = (int)(((float)lane + 0.5f) / (float)rp.nn[1]), iy = lane - iz * rp.nn[1];
		const int wy = wrap_near(rp.lo[1] + iy, g.S[1]), wz = wrap_near(rp.lo[2] + iz, g.S[2]);
		Cell* dst = raw + lane * rp.nn[0];
		int ix = 0;
		while (ix < rp.nn[0]) {   // contiguous run in x up to the periodic boundary
			const int wx = wrap_near(rp.lo[0] + ix, g.S[0]);
			int run = rp.nn[0] - ix; if (run > g.S[0] - wx) run = g.S[0] - wx;
			tma_load_1d(dst + ix, T.cells + cell_index(g, wx, wy, wz), (unsigned)run * (unsigned)sizeof(Cell), mbar);
			ix += run;
		}
	}
}

// Alternative staging with per-lane 16-byte cp.async (LDGSTS): lane handles cells lane, lane+32 (2 chunks of 16 B each).
// Kept because the issue cost of bulk copies matters here: 16 row copies issued by 16 lanes measured ~1.7k cycles
// of issue time per round, against ~0.3k for 16 LDGSTS per lane (profiles/).
__device__ __forceinline__ void region_issue_ldgsts(const Geom& g, const Tables& T, const RegionPlan& rp, Cell* raw, int lane)
{
	const float inv0 = 1.0f / (float)rp.nn[0], inv1 = 1.0f / (float)rp.nn[1];
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int ci = lane + 32 * q;
		if (ci < rp.ncell) {
			const int rest = (int)(((float)ci + 0.5f) * inv0), ix = ci - rest * rp.nn[0];
			const int iz = (int)(((float)rest + 0.5f) * inv1), iy = rest - iz * rp.nn[1];
			const Cell* src = T.cells + cell_index(g, wrap_near(rp.lo[0] + ix, g.S[0]), wrap_near(rp.lo[1] + iy, g.S[1]), wrap_near(rp.lo[2] + iz, g.S[2]));
			const uint32_t dst = smem_u32(raw + ci);
#pragma unroll
			for (int k = 0; k < (int)(sizeof(Cell) / 16); k++)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * k), "l"(reinterpret_cast<const char*>(src) + 16 * k) : "memory");
		}
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void region_wait_ldgsts()
{	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncwarp();
}

// compaction of the staged cells into the warp's neighbour list, positions relative to the (wrapped) last bead. Only beads that
// count (older than this round, not the bonded neighbour) are listed; the overflow-table beads of flagged cells are appended,
// so the list is complete and no later step looks at the table again. Returns the list length or -1 if it does not fit.
__device__ __forceinline__ int region_consume(const Geom& g, const Tables& T, const RegionPlan& rp, const Cell* raw, uint32_t idLimit, uint32_t ignoreId,
	float4* nb, int nbCap, int lane)
{
	const float inv0 = 1.0f / (float)rp.nn[0], inv1 = 1.0f / (float)rp.nn[1];
	uint32_t vm[2] = { 0, 0 }; float sh[2][3]; bool flg[2] = { false, false }; int at_[2] = { 0, 0 };
	uint32_t w[2][CELL_WORDS];
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int ci = lane + 32 * q;
		sh[q][0] = sh[q][1] = sh[q][2] = 0.f;
		if (ci < rp.ncell) {
			const int rest = (int)(((float)ci + 0.5f) * inv0), ix = ci - rest * rp.nn[0];
			const int iz = (int)(((float)rest + 0.5f) * inv1), iy = rest - iz * rp.nn[1];
			at_[q] = region_index(rp, ix, iy, iz);
			const uint4* sp = reinterpret_cast<const uint4*>(raw + at_[q]);   // the record in registers: two 128-bit shared loads
			const uint4 lo = sp[0], hi = sp[1];
			w[q][0] = lo.x; w[q][1] = lo.y; w[q][2] = lo.z; w[q][3] = lo.w; w[q][4] = hi.x; w[q][5] = hi.y; w[q][6] = hi.z; w[q][7] = hi.w;
#pragma unroll
			for (int k = 0; k < MAX_CAP; k++) { const uint32_t id = w[q][k]; vm[q] |= (id < idLimit && id != ignoreId ? 1u : 0u) << k; }
			flg[q] = cell_flagged(w[q]);
			sh[q][0] = rp.base[0] + (float)ix * (float)g.a[0]; sh[q][1] = rp.base[1] + (float)iy * (float)g.a[1]; sh[q][2] = rp.base[2] + (float)iz * (float)g.a[2];
		}
	}
	const uint32_t mine = (uint32_t)(__popc(vm[0]) + __popc(vm[1]));                 // <= 6
	const unsigned lt = (1u << lane) - 1u;
	uint32_t excl = 0, total = 0;
#pragma unroll
	for (int b = 0; b < 3; b++) { const unsigned m = __ballot_sync(FULL, (mine >> b) & 1u); excl += (uint32_t)__popc(m & lt) << b; total += (uint32_t)__popc(m) << b; }
	unsigned flagged = __ballot_sync(FULL, flg[0] || flg[1]);
	if (total > (uint32_t)nbCap) return -1;
	uint32_t at = excl;
#pragma unroll
	for (int q = 0; q < 2; q++) {
#pragma unroll
		for (int k = 0; k < MAX_CAP; k++)
			if ((vm[q] >> k) & 1u)
				nb[at++] = make_float4(sh[q][0] + dec_off(g, cell_off(w[q], k, 0), 0), sh[q][1] + dec_off(g, cell_off(w[q], k, 1), 1),
					sh[q][2] + dec_off(g, cell_off(w[q], k, 2), 2), __uint_as_float(w[q][k]));
	}
	while (flagged) {   // rare: one lane at a time walks the hash table for its flagged cells and appends what it finds
		const int src = __ffs(flagged) - 1;
		uint32_t add = 0;
		if (lane == src) {
#pragma unroll 1
			for (int q = 0; q < 2; q++) {
				if (!flg[q]) continue;
				const int ci = lane + 32 * q;
				const int rest = (int)(((float)ci + 0.5f) * inv0), ix = ci - rest * rp.nn[0];
				const int iz = (int)(((float)rest + 0.5f) * inv1), iy = rest - iz * rp.nn[1];
				const int64_t cidx = cell_index(g, wrap_near(rp.lo[0] + ix, g.S[0]), wrap_near(rp.lo[1] + iy, g.S[1]), wrap_near(rp.lo[2] + iz, g.S[2]));
				ovf_for_cell(T, cidx, [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
					if (id < idLimit && id != ignoreId) {
						if (total + add < (uint32_t)nbCap)
							nb[total + add] = make_float4(sh[q][0] + dec_off(g, o0, 0), sh[q][1] + dec_off(g, o1, 1), sh[q][2] + dec_off(g, o2, 2), __uint_as_float(id));
						add++;
					}
					return false;
				});
			}
		}
		total += __shfl_sync(FULL, add, src);
		flagged &= flagged - 1u;
	}
	__syncwarp();
	if (total > (uint32_t)nbCap) return -1;
	return (int)total;
}

// One chain, one round, one warp: returns the accepted trial (1-based) or 0, with the bead in cand / vAcc (fractional).
// No global writes except marking a chain dead. Fast path: the exact candidates of trials 1..32 are prepared lane-parallel
// while the region is being staged, then tested lane-per-trial; chains that found nothing there (`hard`) go through the
// blocked-arc classification of all remaining trials.
// -DSARW_PROF: cycles between the steps of the fast path, summed over every chain-round (slot k), and a histogram of the whole
// propose time per chain-round in 512-cycle bins (slots 16..47); read back with sarw_debug_prof. Off in the shipped library.
#ifdef SARW_PROF
constexpr int PROF_ROWS = 4096;                         // one row of 64 counters per warp (mod 4096): no contention between warps
__device__ unsigned long long g_prof[64 * PROF_ROWS];
#define PROF_ROW() (g_prof + 64 * ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (PROF_ROWS - 1)))
#define PROF_BEGIN() long long prof_t0_ = clock64(), prof_tp_ = prof_t0_
#define PROF_MARK(k) do { const long long t_ = clock64(); if (lane == 0) atomicAdd(PROF_ROW() + (k), (unsigned long long)(t_ - prof_tp_)); prof_tp_ = t_; } while (0)
#define PROF_END() do { if (lane == 0) { long long d_ = (clock64() - prof_t0_) >> 9; d_ = d_ > 31 ? 31 : d_; atomicAdd(PROF_ROW() + 16 + d_, 1ull); atomicAdd(PROF_ROW() + 15, 1ull); } } while (0)
#else
#define PROF_BEGIN() do { } while (0)
#define PROF_MARK(k) do { } while (0)
#define PROF_END() do { } while (0)
#endif

__device__ __forceinline__ uint32_t propose_chain(const Geom& g, const Tables& T, const void* tmap, Control* ctl, ChainState* chains, uint32_t i, uint32_t round,
	uint32_t idRound, ChainState& cs, const Frame& f, const RegionPlan& rp, bool hard, Cell* raw, float4* nb, int nbCap, float4* arcs, int arcCap, void* mbar,
	unsigned& tmaParity, int lane, double cand[3], double vAcc[3], unsigned long long& queries)
{
	const uint32_t T_ = (uint32_t)g.maxTrials;
	uint32_t accepted = 0;
	int nnb = -1; int narcs = -1;
	const bool canStage = rp.ncell <= CL_REGION_CELLS;
	PROF_BEGIN();
#ifndef SARW_STAGE_LDGSTS
	if (canStage) region_issue_tma(g, T, tmap, rp, raw, mbar, lane);   // bulk copies in flight ...
#else
	if (canStage) region_issue_ldgsts(g, T, rp, raw, lane);      // async copies in flight ...
#endif
	PROF_MARK(0);   // copies issued
	// ... while every lane prepares the exact candidate of trial lane+1 (reference arithmetic)
	double candL[3], vL[3]; bool okL = false; float rel[3] = { 0, 0, 0 };
	if (!hard) {
		const uint4 w = philox4x32_10(make_uint4(i, round, (uint32_t)lane >> 2, 0u), g.key0, g.key1);
		cone_candidate(g, f, cs.last, (double)pick(w, (uint32_t)lane & 3u) * (1.0 / 4294967296.0), candL);
		okL = (uint32_t)lane < T_ && !growth_rules_reject(g, candL, cs.last);
		to_frac(g, candL, vL);
		rel[0] = (float)(candL[0] - cs.last[0]); rel[1] = (float)(candL[1] - cs.last[1]); rel[2] = (float)(candL[2] - cs.last[2]);
	}
	PROF_MARK(1);   // candidates of trials 1..32 computed
	if (canStage) {
#ifndef SARW_STAGE_LDGSTS
		mbar_wait(mbar, tmaParity); tmaParity ^= 1u;
#else
		region_wait_ldgsts();
#endif
		PROF_MARK(2);   // waited for the copies
		nnb = region_consume(g, T, rp, raw, idRound, cs.lastId, nb, nbCap, lane);
		PROF_MARK(3);   // neighbour list compacted
#ifdef SARW_PROF
		if (lane == 0) { atomicAdd(PROF_ROW() + 8, (unsigned long long)(nnb > 0 ? nnb : 0)); atomicAdd(PROF_ROW() + 9, (unsigned long long)rp.ncell); atomicAdd(PROF_ROW() + 10, (unsigned long long)(rp.nn[1] * rp.nn[2])); }
#endif
	}
#ifdef SARW_PROF
	PROF_MARK(5);
#endif
	if (!hard) {
		// fast path: lane-per-trial exact test of trials 1..32 against the staged neighbourhood
		if (okL) {
			if (nnb >= 0) {
				// branch-free pass over the staged beads: smallest squared distance and the bead that attains it
				float dmin = 3.0e38f; int jmin = -1; int inBand = 0;
#pragma unroll 4
				for (int j = 0; j < nnb; ++j) {
					const float4 b = nb[j];
					const float dx = rel[0] - b.x, dy = rel[1] - b.y, dz = rel[2] - b.z;
					const float d2 = dx * dx + dy * dy + dz * dz;
					inBand += (d2 <= g.hi2);
					if (d2 < dmin) { dmin = d2; jmin = j; }
				}
#ifdef SARW_PROF
				{ const unsigned anyBand = __ballot_sync(__activemask(), inBand && !(dmin < g.lo2)); if (anyBand && lane == (__ffs(__activemask()) - 1)) atomicAdd(PROF_ROW() + 11, 1ull); }
#endif
				if (dmin < g.lo2) okL = false;
				else if (inBand) {   // some bead within the undecided band (rare): exact reference arithmetic on those
					for (int j = 0; j < nnb && okL; ++j) {
						const float4 b = nb[j];
						const float dx = rel[0] - b.x, dy = rel[1] - b.y, dz = rel[2] - b.z;
						const float d2 = dx * dx + dy * dy + dz * dz;
						if (d2 <= g.hi2 && exact_pair(g, vL, T.lpos + 3 * (size_t)__float_as_uint(b.w))) okL = false;
					}
				}
				(void)jmin;
			} else okL = !probe_cells<1>(g, T, vL, 0u, idRound, cs.lastId, 0);
		}
		const unsigned m = __ballot_sync(FULL, okL);
		PROF_MARK(4);   // 32 trials tested
		queries += (lane == 0) ? (T_ < 32u ? T_ : 32u) : 0u;
		if (m) {
			const int first = __ffs(m) - 1;
			accepted = (uint32_t)first + 1u;
			for (int k = 0; k < 3; k++) { cand[k] = __shfl_sync(FULL, candL[k], first); vAcc[k] = __shfl_sync(FULL, vL[k], first); }
		}
	}
	if (!accepted && T_ > (hard ? 0u : 32u)) {
		// hard chain: blocked-arc classification of the remaining trials, exact evaluation of the survivors
		narcs = nnb >= 0 ? build_arcs_cs(g, f, cs.last, nb, nnb, arcs, arcCap, lane) : -1;
		if (lane == 0 && nnb < 0) atomicAdd(&ctl->stageFallbacks, 1ull);
		if (lane == 0 && nnb >= 0 && narcs < 0) atomicAdd(&ctl->arcFallbacks, 1ull);
		const uint32_t skip = hard ? 0u : 32u;          // trials <= skip were just evaluated exactly and failed
		for (uint32_t base = 0; base < T_ && !accepted; base += 128) {
			const uint32_t blk = (base >> 2) + lane;
			const uint4 w = philox4x32_10(make_uint4(i, round, blk, 0u), g.key0, g.key1);
			const uint32_t wj[4] = { w.x, w.y, w.z, w.w };
			unsigned cm[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t t = base + 4 * lane + j + 1;
				bool candidate = t <= T_ && t > skip;
				if (candidate && narcs > 0) {
					const float theta = (float)wj[j] * (float)(359.0 * 3.14159265358979323846 / 180 / 4294967296.0);
					float s, c;
					__sincosf(theta - 3.14159265358979323846f, &s, &c);   // |arg| <= pi: fast intrinsic is accurate to ~1e-6
					s = -s; c = -c;
					for (int a = 0; a < narcs; ++a) {
						const float4 arc = arcs[a];
						if (arc.x * c + arc.y * s >= arc.z) { candidate = false; break; }
					}
				}
				cm[j] = __ballot_sync(FULL, candidate);
			}
			unsigned any = cm[0] | cm[1] | cm[2] | cm[3];
			while (any) {   // candidates in increasing trial order
				const int l = __ffs(any) - 1;
				int j = 0;
#pragma unroll
				for (int jj = 3; jj >= 0; --jj) if ((cm[jj] >> l) & 1u) j = jj;
				const uint32_t word = __shfl_sync(FULL, wj[0], l) * (j == 0) + __shfl_sync(FULL, wj[1], l) * (j == 1)
					+ __shfl_sync(FULL, wj[2], l) * (j == 2) + __shfl_sync(FULL, wj[3], l) * (j == 3);
				queries += (lane == 0);
				if (eval_trial_coop(g, T, f, cs, (double)word * (1.0 / 4294967296.0), nb, nnb, idRound, lane, cand)) {
					accepted = base + 4 * l + j + 1;
					to_frac(g, cand, vAcc);
					break;
				}
#pragma unroll
				for (int jj = 0; jj < 4; ++jj) if (jj == j) cm[jj] &= ~(1u << l);
				any = cm[0] | cm[1] | cm[2] | cm[3];
			}
		}
		if (!accepted && narcs > 0 && arcs_cs_cover_circle(arcs, narcs, reinterpret_cast<double2*>(nb), lane)) {   // nb is no longer needed
			cs.flags |= CHAIN_DEAD;
			if (lane == 0) { chains[i].flags = cs.flags; atomicAdd(&ctl->deadChains, 1u); }
		}
	}
	PROF_END();
	return accepted;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) grow_cluster_kernel(const __grid_constant__ GrowArgs A)
{
	extern __shared__ __align__(128) unsigned char clSmem[];
	Cell* rawAll = reinterpret_cast<Cell*>(clSmem);
	constexpr int NBCAP = cl_nb_cap(WARPS), ARCCAP = cl_arc_cap(WARPS);
	float4* nbAll = reinterpret_cast<float4*>(rawAll + WARPS * CL_REGION_CELLS);
	float4* arcAll = nbAll + WARPS * NBCAP;
	TentRec* table = reinterpret_cast<TentRec*>(arcAll + WARPS * ARCCAP);
	unsigned long long* mbarAll = reinterpret_cast<unsigned long long*>(table + 2 * CL_MAXCH);
	volatile uint32_t* dirtyWord = reinterpret_cast<volatile uint32_t*>(mbarAll + WARPS);

	cgx::cluster_group cluster = cgx::this_cluster();
	const Geom& g = A.g;
	const unsigned nCta = cluster.num_blocks(), myCta = cluster.block_rank();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t i = myCta * WARPS + warp;            // the chain this warp owns for the whole launch
	const bool owner = i < g.nChains;
	Cell* raw = rawAll + warp * CL_REGION_CELLS;
	float4* nb = nbAll + warp * NBCAP;
	float4* arcs = arcAll + warp * ARCCAP;
	void* mbar = mbarAll + warp;
	const uint32_t T_ = (uint32_t)g.maxTrials;

	if (threadIdx.x == 0) { dirtyWord[0] = 0; dirtyWord[1] = 0; }
	if (lane == 0) mbar_init(mbar, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	ChainState cs;
	if (owner) cs = load_chain(A.chains + i); else memset(&cs, 0, sizeof cs);
	Frame f; RegionPlan rp;
	memset(&f, 0, sizeof f); memset(&rp, 0, sizeof rp);
	if (owner && cs.lastId != ID_NONE) prepare_chain(g, cs, f, rp);
	bool hard = false;                                   // no trial among the first 32 passed last round: go straight to the arcs
	unsigned tmaParity = 0;
	cluster_arrive(); cluster_wait();                    // every CTA is running and its shared memory is initialised

	uint32_t round = __ldcg(&A.ctl->round);
	unsigned long long nBeads = __ldcg(&A.ctl->nBeads);
	uint32_t lastProg = __ldcg(&A.ctl->lastProgressed);
	const bool go = __ldcg(&A.ctl->initiated) != 0u;   // sarw.cpp:66 (see grow_kernel)
	unsigned long long myTrials = 0, myQueries = 0; long long myBusy = 0;
	const bool timer = (myCta == 0 && threadIdx.x == 0);
	long long cycA = 0, cycV = 0, cycC = 0; uint32_t cleanRounds = 0;
	long long sub[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, ts0 = 0, ts1 = 0;
	long long pathCyc[4] = { 0, 0, 0, 0 }; uint32_t pathCnt[4] = { 0, 0, 0, 0 };
#define SUBT(k) do { ts1 = clock64(); sub[k] += ts1 - ts0; ts0 = ts1; } while (0)

	while (go && round < A.roundEnd && (long long)nBeads < g.target && lastProg) {
		const long long tk0 = clock64();
		ts0 = tk0;
		uint32_t start, end;
		round_range(g, nBeads, start, end);
		const uint32_t idRound = round_id_base(g, round);
		const int par = (int)(round & 1u);
		const bool active = owner && i >= start && i < end;

		// ---------------- phase A: no global writes
		uint32_t accepted = 0; double cand[3] = { 0, 0, 0 };
		TentRec rec; rec.x = rec.y = rec.z = 0; rec.trial = 0; rec.cx = rec.cy = rec.cz = 0;
		uint32_t q16[3] = { 0, 0, 0 };                     // quantised in-cell offset of the accepted bead (for the slot store)
		double vAcc[3] = { 0, 0, 0 };                     // its fractional coordinates
		uint32_t freeMask = 0x80000000u;                   // free slots of its cell at round start (from the staged copy); top bit = unknown
		int path = 0;
		if (active && (cs.flags & CHAIN_DEAD)) rec.trial = T_ + 1u;
		else if (active) {
			path = hard ? 3 : 1;
			accepted = propose_chain(g, A.T, A.tmap, A.ctl, A.chains, i, round, idRound, cs, f, rp, hard, raw, nb, NBCAP, arcs, ARCCAP, mbar, tmaParity, lane, cand, vAcc, myQueries);
			hard = accepted == 0u || accepted > 32u;
			if (path == 1 && hard) path = 2;
			rec.trial = accepted ? accepted : T_ + 1u;
			if (accepted) {
				rec.x = cand[0]; rec.y = cand[1]; rec.z = cand[2];
				int cc[3];
				bead_cell(g, vAcc, q16, cc);
				rec.cx = cc[0]; rec.cy = cc[1]; rec.cz = cc[2];
				if (rp.ncell <= CL_REGION_CELLS) {   // free slots of the bead's cell at round start, from the staged copy (the cell lies inside the region)
					const int ix = wrap_near(cc[0] - rp.lo[0], g.S[0]), iy = wrap_near(cc[1] - rp.lo[1], g.S[1]), iz = wrap_near(cc[2] - rp.lo[2], g.S[2]);
					if (ix < rp.nn[0] && iy < rp.nn[1] && iz < rp.nn[2]) {
						const Cell& sc = raw[region_index(rp, ix, iy, iz)];
						freeMask = (sc.id[0] == ID_NONE ? 1u : 0u) | (sc.id[1] == ID_NONE ? 2u : 0u) | (sc.id[2] == ID_NONE ? 4u : 0u);
					}
				}
			}
			SUBT(3);
		}
		// broadcast the record into every CTA's table (distributed shared memory)
		if (owner)
			for (unsigned c = lane; c < nCta; c += 32) *cluster.map_shared_rank(&table[par * CL_MAXCH + i], c) = rec;
		SUBT(4);
		myBusy += clock64() - tk0;
		if (active) { pathCyc[path] += clock64() - tk0; pathCnt[path]++; }
		cluster_arrive();
		// in the shadow of the barrier: next round's chain state, frame and region, assuming the tentative bead stands (it almost
		// always does). The old state is recoverable: old last = new pen, old lastId = new penId; only the old pen is kept aside.
		double oldPen[3] = { cs.pen[0], cs.pen[1], cs.pen[2] }; const uint32_t oldPenId = cs.penId;
		const bool advanced = active && accepted;
		if (advanced) {
			for (int k = 0; k < 3; k++) { cs.pen[k] = cs.last[k]; cs.last[k] = cand[k]; }
			cs.penId = cs.lastId; cs.lastId = idRound + i; cs.length += 1;
			prepare_chain(g, cs, f, rp);
		}
		SUBT(5);
		cluster_wait();
		SUBT(6);
		if (threadIdx.x == 0) dirtyWord[par ^ 1] = 0;    // last round's flag has been read by every warp of this CTA
		const long long tk1 = clock64();

		// ---------------- phase V from the local table: round totals, conflicts with lower chains, slot rank
		uint32_t acc = 0, prog = 0, rank = 0; bool conflict = false;
		{
			// two records per lane and iteration, the small half of a record first (trial + cell): the position is read only for the rare
			// record in an adjacent cell. Independent loads in flight instead of one dependent 40-byte read per iteration.
			for (uint32_t jb = start; jb < end; jb += 64) {
				uint32_t jj[2] = { jb + (uint32_t)lane, jb + 32u + (uint32_t)lane };
				uint2 tc[2], cyz[2];
#pragma unroll
				for (int u = 0; u < 2; u++) {
					const TentRec* rp_ = &table[par * CL_MAXCH + (jj[u] < end ? jj[u] : start)];
					tc[u] = *reinterpret_cast<const uint2*>(&rp_->trial);      // {trial, cx}
					cyz[u] = *reinterpret_cast<const uint2*>(&rp_->cy);        // {cy, cz}
				}
#pragma unroll
				for (int u = 0; u < 2; u++) {
					const uint32_t j = jj[u];
					if (j >= end) continue;
					const uint32_t trial = tc[u].x;
					const int rcx = (int)tc[u].y, rcy = (int)cyz[u].x, rcz = (int)cyz[u].y;
					const bool has = trial >= 1u && trial <= T_;
					acc += has; prog += (has && trial < T_);
					if (!accepted || !has || j == i) continue;
					int dx = abs(rcx - rec.cx), dy = abs(rcy - rec.cy), dz = abs(rcz - rec.cz);
					if (dx == 0 && dy == 0 && dz == 0) rank += (j < i);
					if (j < i) {   // only beads in the same or an adjacent cell (periodic) can be within thresh
						dx = min(dx, g.S[0] - dx); dy = min(dy, g.S[1] - dy); dz = min(dz, g.S[2] - dz);
						if (dx <= 1 && dy <= 1 && dz <= 1) {
							const TentRec& r = table[par * CL_MAXCH + j];
							const double pj[3] = { r.x, r.y, r.z };
							conflict = conflict || exact_pair_reg(g, vAcc, pj);
						}
					}
				}
			}
			// one hardware warp reduction over packed counters (each <= 256 chains: 10 bits are plenty)
			const unsigned packed = __reduce_add_sync(FULL, acc | (prog << 10) | (rank << 20));
			acc = packed & 1023u; prog = (packed >> 10) & 1023u; rank = packed >> 20;
			conflict = __any_sync(FULL, conflict);
		}
		SUBT(0);   // table scan + reductions
		// ---------------- insert with plain stores: slot = count seen while staging + rank among this round's beads of the cell
		if (active) {
			if (accepted && lane == 0) {
				const uint32_t id = idRound + i;
				const int64_t cidx = cell_index(g, rec.cx, rec.cy, rec.cz);
				Cell* cell = A.T.cells + cidx;
				uint32_t fm = freeMask;
				if (fm & 0x80000000u) {   // the cell was not staged (crowded region): its ids as of round start from global memory
					const uint4 w = __ldcg(reinterpret_cast<const uint4*>(cell));
					fm = (w.x == ID_NONE ? 1u : 0u) | (w.y == ID_NONE ? 2u : 0u) | (w.z == ID_NONE ? 4u : 0u);
				}
				fm &= (1u << g.cap) - 1u;
				double* lp = A.T.lpos + 3 * (size_t)id;
				lp[0] = cand[0]; lp[1] = cand[1]; lp[2] = cand[2];
				for (uint32_t r = 0; r < rank && fm; ++r) fm &= fm - 1u;   // the rank-th free slot: every warp derives the same assignment
				if (fm) {
					const int k = __ffs((int)fm) - 1;
					st_off16(cell, k, q16);
					__stcg(&cell->id[k], id);
				} else {           // cell full: overflow table (rare)
					cell->ovfInv = 0;
					ovf_insert(A.T, cidx, id, q16);
				}
				fence_cells_for_tma();
				for (int k = 0; k < 3; k++) A.resPos[3 * (size_t)i + k] = cand[k];
			}
			if (lane == 0) {
				A.resTrial[i] = rec.trial; A.origTrial[i] = rec.trial;
				myTrials += rec.trial <= T_ ? rec.trial : T_;
				if (conflict) { const uint32_t at = atomicAdd(&A.dirtyCount[round], 1u); A.dirtyList[at] = i; }
			}
			if (conflict)
				for (unsigned c = lane; c < nCta; c += 32) *cluster.map_shared_rank(const_cast<uint32_t*>(&dirtyWord[par]), c) = 1u;
		}
		if (i == 0 && lane == 0) { A.roundAcc[round] = acc; A.roundProg[round] = prog; }
		SUBT(1);   // insert + result stores
		cluster_arrive(); cluster_wait();
		SUBT(2);   // second barrier
		const long long tk2 = clock64();

		// ---------------- clean-up (rare): the shared serial path through global memory
		if (dirtyWord[par] != 0u) {
			// the clean-up reads chain states from global memory as of round start: they are (commits happen below)
			if (i == 0) cleanup_round(A, round, lane);
			cluster_arrive(); cluster_wait();
			cleanRounds++;
			if (active) {
				const uint32_t t = __ldcg(&A.resTrial[i]);
				const uint32_t nowAcc = (t >= 1u && t <= T_) ? t : 0u;
				double np[3] = { 0, 0, 0 };
				if (nowAcc) for (int k = 0; k < 3; k++) np[k] = __ldcg(&A.resPos[3 * (size_t)i + k]);
				const bool same = (nowAcc != 0u) == advanced && (!nowAcc || (np[0] == cand[0] && np[1] == cand[1] && np[2] == cand[2]));
				if (!same) {   // roll the speculative state back to round start, then apply what the clean-up decided
					if (advanced) {
						for (int k = 0; k < 3; k++) { cs.last[k] = cs.pen[k]; cs.pen[k] = oldPen[k]; }
						cs.lastId = cs.penId; cs.penId = oldPenId; cs.length -= 1;
					}
					if (nowAcc) {
						for (int k = 0; k < 3; k++) { cs.pen[k] = cs.last[k]; cs.last[k] = np[k]; }
						cs.penId = cs.lastId; cs.lastId = idRound + i; cs.length += 1;
					}
					prepare_chain(g, cs, f, rp);
				}
				accepted = nowAcc;
				hard = hard || accepted == 0u;
			}
			acc = __ldcg(&A.roundAcc[round]); prog = __ldcg(&A.roundProg[round]);
		}
		// ---------------- publish the committed chain state (sarw.cpp:270-273)
		if (active && lane == 0) {
			if (accepted) { A.prev[idRound + i - g.idAtomBase] = cs.penId; A.chains[i] = cs; }
			A.resTrial[i] = 0;
		}
		nBeads += acc; lastProg = prog != 0u; round++;
		cycA += tk1 - tk0; cycV += tk2 - tk1; cycC += clock64() - tk2;
	}
#undef SUBT
	if (lane == 0) {
		if (myTrials) atomicAdd(&A.ctl->trials, myTrials);
		if (myQueries) atomicAdd(&A.ctl->queries, myQueries);
	}
	if (i == 0 && lane == 0) { A.ctl->round = round; A.ctl->nBeads = nBeads; A.ctl->lastProgressed = lastProg ? 1u : 0u; }
	if (timer) for (int k = 0; k < 8; k++) A.ctl->cycSub[k] += sub[k];
	if (lane == 0 && i < 256) A.ctl->busy[i] += (unsigned long long)myBusy;
	if (lane == 0 && owner) for (int k = 0; k < 4; k++) { atomicAdd(&A.ctl->pathCyc[k], (unsigned long long)pathCyc[k]); atomicAdd(&A.ctl->pathCnt[k], (unsigned long long)pathCnt[k]); }
	if (timer) { A.ctl->cycPropose += cycA; A.ctl->cycVerify += cycV; A.ctl->cycCleanup += cycC; A.ctl->cleanupRounds += cleanRounds; }
	cluster_arrive(); cluster_wait();                    // no CTA may exit while others can still write into its shared memory
}
