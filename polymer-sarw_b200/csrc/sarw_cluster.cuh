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
//   * beads enter the cell list AFTER that barrier with plain stores: the slot is old count (seen while staging) + rank
//     among this round's beads of the same cell, which every warp derives from the same table — no atomic round trip;
//   * the second cluster barrier publishes the stores and the (rare) conflict flag; conflicts fall back to the shared
//     serial clean-up (cleanup_round) through global memory.
#pragma once
// (sarw_kernels.cu includes <cooperative_groups.h> at file scope before opening namespace sarw)
namespace cgx = ::cooperative_groups;

constexpr int CL_WARPS = 16;            // warps (= chains) per CTA
constexpr int CL_MAX_CTAS = 16;         // non-portable cluster size
constexpr int CL_MAXCH = CL_WARPS * CL_MAX_CTAS;
constexpr int CL_NARC = 64;             // blocked arcs kept per chain (more: no pre-classification, still exact)

// tentative bead of one chain in one round, as broadcast to every CTA
struct __align__(8) TentRec {
	double x, y, z;
	uint32_t trial;     // 0 = chain not in this round's range, 1..T accepted trial, T+1 none
	int32_t cx, cy, cz; // wrapped cell coordinates of the bead
};
static_assert(sizeof(TentRec) == 40, "TentRec is 40 bytes");

constexpr size_t cluster_smem_bytes()
{	return (size_t)CL_WARPS * NBMAX * sizeof(float4) + (size_t)CL_WARPS * CL_NARC * sizeof(float4) + 2 * (size_t)CL_MAXCH * sizeof(TentRec) + 16;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Blocked arcs as (cos alpha, sin alpha, kappa): a trial angle theta is CERTAINLY rejected by the bead iff
// cos(alpha) cos(theta) + sin(alpha) sin(theta) >= kappa. kappa carries margins (1e-5 A in distance, 3e-5 in cos) that
// dominate the fp32 evaluation error of both sides; arcs only ever skip trials, the exact test decides everything else.
__device__ __forceinline__ int build_arcs_cs(const Geom& g, const Frame& f, const float4* nb, int nnb, float4* arcs, int lane)
{
	const double rho = g.dSinPhi, h = g.dCosPhi;
	const double ts = g.thresh - 1e-5;
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
		if (narcs + __popc(m) > CL_NARC) return -1;
		if (have) arcs[narcs + __popc(m & ((1u << lane) - 1u))] = arc;
		narcs += __popc(m);
	}
	__syncwarp();
	return narcs;
}

// Same coverage criterion as arcs_cover_circle, evaluated in double from the (cos, sin, kappa) form (rare path).
__device__ __forceinline__ bool arcs_cs_cover_circle(const float4* arcs, int narcs, int lane)
{
	if (narcs <= 0) return false;
	const double PI = 3.14159265358979323846;
	const double gapLo = 359.0 * PI / 180 - 1e-9, gapHi = 2 * PI;
	auto centre = [&](int i) { return i < narcs ? atan2((double)arcs[i].y, (double)arcs[i].x) : 0.5 * (gapLo + gapHi); };
	auto width = [&](int i) {
		if (i >= narcs) return 0.5 * (gapHi - gapLo) - 1e-7;
		const double k = (double)arcs[i].z;
		return k <= -1.0 ? PI : (k >= 1.0 ? 0.0 : acos(k) - 1e-6);
	};
	bool allCovered = true;
	for (int ib = 0; ib <= narcs; ib += 32) {
		const int i = ib + lane;
		bool covered = true;
		if (i <= narcs) {
			const double e = centre(i) + width(i);
			covered = false;
			for (int j = 0; j <= narcs && !covered; ++j) {
				const double wj = width(j);
				if (wj >= PI) { covered = true; break; }
				double d = e - centre(j);
				d -= 2 * PI * rint(d / (2 * PI));
				covered = fabs(d) < wj - 1e-6;
			}
		}
		allCovered = allCovered && __all_sync(FULL, covered);
	}
	return allCovered;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) grow_cluster_kernel(GrowArgs A)
{
	extern __shared__ __align__(16) unsigned char clSmem[];
	float4* nbAll = reinterpret_cast<float4*>(clSmem);
	float4* arcAll = nbAll + WARPS * NBMAX;
	TentRec* table = reinterpret_cast<TentRec*>(arcAll + WARPS * CL_NARC);
	volatile uint32_t* dirtyWord = reinterpret_cast<volatile uint32_t*>(table + 2 * CL_MAXCH);

	cgx::cluster_group cluster = cgx::this_cluster();
	const Geom& g = A.g;
	const unsigned nCta = cluster.num_blocks(), myCta = cluster.block_rank();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t i = myCta * WARPS + warp;            // the chain this warp owns for the whole launch
	const bool owner = i < g.nChains;
	float4* nb = nbAll + warp * NBMAX;
	float4* arcs = arcAll + warp * CL_NARC;
	const uint32_t T_ = (uint32_t)g.maxTrials;

	if (threadIdx.x == 0) { dirtyWord[0] = 0; dirtyWord[1] = 0; }
	ChainState cs;
	if (owner) cs = load_chain(A.chains + i); else memset(&cs, 0, sizeof cs);
	cluster_arrive(); cluster_wait();                    // every CTA is running and its shared memory is initialised

	uint32_t round = __ldcg(&A.ctl->round);
	unsigned long long nBeads = __ldcg(&A.ctl->nBeads);
	uint32_t lastProg = __ldcg(&A.ctl->lastProgressed);
	unsigned long long myTrials = 0, myQueries = 0;
	const bool timer = (myCta == 0 && threadIdx.x == 0);
	long long cycA = 0, cycV = 0, cycC = 0; uint32_t cleanRounds = 0;

	while (round < A.roundEnd && (long long)nBeads < g.target && lastProg) {
		const long long tk0 = clock64();
		uint32_t start, end;
		round_range(g, nBeads, start, end);
		const uint32_t idRound = round_id_base(g, round);
		const int par = (int)(round & 1u);
		const bool active = owner && i >= start && i < end;

		// ---------------- phase A: stage, classify, evaluate — no global writes
		uint32_t accepted = 0; double cand[3] = { 0, 0, 0 };
		int64_t stagedCell[2] = { -1, -1 }; uint32_t stagedCnt[2] = { 0, 0 };   // cells this lane staged (for the slot computation)
		TentRec rec; rec.x = rec.y = rec.z = 0; rec.trial = 0; rec.cx = rec.cy = rec.cz = 0;
		if (active && (cs.flags & CHAIN_DEAD)) rec.trial = T_ + 1u;
		else if (active) {
			Frame f; cone_frame(cs.last, cs.pen, f);
			double lw[3];
			{ double vl[3]; to_frac(g, cs.last, vl); for (int k = 0; k < 3; k++) lw[k] = vl[k] * g.L[k]; }
			bool regionOvf = false;
			const int nnb = stage_region(g, A.T, lw, f, idRound, cs.lastId, nb, lane, regionOvf, stagedCell, stagedCnt);
			const int narcs = nnb >= 0 ? build_arcs_cs(g, f, nb, nnb, arcs, lane) : -1;
			const uint32_t novf = regionOvf ? __ldcg(A.T.ovfCount) : 0u;
			for (uint32_t base = 0; base < T_ && !accepted; base += 128) {
				const uint32_t blk = (base >> 2) + lane;
				const uint4 w = philox4x32_10(make_uint4(i, round, blk, 0u), g.key0, g.key1);
				const uint32_t wj[4] = { w.x, w.y, w.z, w.w };
				unsigned cm[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t t = base + 4 * lane + j + 1;
					bool candidate = t <= T_;
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
					myQueries += (lane == 0);
					if (eval_trial_coop(g, A.T, f, cs, (double)word * (1.0 / 4294967296.0), nb, nnb, idRound, novf, lane, cand)) {
						accepted = base + 4 * l + j + 1;
						break;
					}
#pragma unroll
					for (int jj = 0; jj < 4; ++jj) if (jj == j) cm[jj] &= ~(1u << l);
					any = cm[0] | cm[1] | cm[2] | cm[3];
				}
			}
			if (!accepted && narcs > 0 && arcs_cs_cover_circle(arcs, narcs, lane)) {
				cs.flags |= CHAIN_DEAD;
				if (lane == 0) { A.chains[i].flags = cs.flags; atomicAdd(&A.ctl->deadChains, 1u); }
			}
			rec.trial = accepted ? accepted : T_ + 1u;
			if (accepted) {
				double v[3];
				to_frac(g, cand, v);
				rec.x = cand[0]; rec.y = cand[1]; rec.z = cand[2];
#pragma unroll
				for (int k = 0; k < 3; k++) {
					int ci = (int)floor(v[k] * g.sOverL[k]);
					if (ci >= g.S[k]) ci = g.S[k] - 1;
					if (ci < 0) ci = 0;
					(k == 0 ? rec.cx : (k == 1 ? rec.cy : rec.cz)) = ci;
				}
			}
		}
		// broadcast the record into every CTA's table (distributed shared memory)
		if (owner)
			for (unsigned c = lane; c < nCta; c += 32) *cluster.map_shared_rank(&table[par * CL_MAXCH + i], c) = rec;
		cluster_arrive(); cluster_wait();
		if (threadIdx.x == 0) dirtyWord[par ^ 1] = 0;    // last round's flag has been read by every warp of this CTA
		const long long tk1 = clock64();

		// ---------------- phase V from the local table: round totals, conflicts with lower chains, slot rank
		uint32_t acc = 0, prog = 0, rank = 0, tot = 0; bool conflict = false;
		{
			double vi[3] = { 0, 0, 0 };
			if (accepted) to_frac(g, cand, vi);
			for (uint32_t jb = start; jb < end; jb += 32) {
				const uint32_t j = jb + lane;
				if (j >= end) continue;
				const TentRec r = table[par * CL_MAXCH + j];
				const bool has = r.trial >= 1u && r.trial <= T_;
				acc += has; prog += (has && r.trial < T_);
				if (!accepted || !has || j == i) continue;
				int dx = abs(r.cx - rec.cx), dy = abs(r.cy - rec.cy), dz = abs(r.cz - rec.cz);
				if (dx == 0 && dy == 0 && dz == 0) { tot++; rank += (j < i); }
				if (j < i) {   // only beads in the same or an adjacent cell (periodic) can be within thresh
					dx = min(dx, g.S[0] - dx); dy = min(dy, g.S[1] - dy); dz = min(dz, g.S[2] - dz);
					if (dx <= 1 && dy <= 1 && dz <= 1) { const double pj[3] = { r.x, r.y, r.z }; conflict = conflict || exact_pair_reg(g, vi, pj); }
				}
			}
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				acc += __shfl_xor_sync(FULL, acc, d); prog += __shfl_xor_sync(FULL, prog, d);
				rank += __shfl_xor_sync(FULL, rank, d); tot += __shfl_xor_sync(FULL, tot, d);
			}
			conflict = __any_sync(FULL, conflict);
		}
		// ---------------- insert with plain stores: slot = count seen while staging + rank among this round's beads of the cell
		if (active) {
			if (accepted) {
				const uint32_t id = idRound + i;
				const int64_t cellIdx = cell_index(g, rec.cx, rec.cy, rec.cz);
				uint32_t old = 0xFFFFFFFFu;
				{	// the bead's cell is inside the staged region: take its count from the lane that staged it
					const bool m0 = stagedCell[0] == cellIdx, m1 = stagedCell[1] == cellIdx;
					const unsigned bm = __ballot_sync(FULL, m0 || m1);
					if (bm) { const int src = __ffs(bm) - 1; const uint32_t mine = m0 ? stagedCnt[0] : stagedCnt[1]; old = __shfl_sync(FULL, mine, src); }
				}
				if (lane == 0) {
					Cell* cell = A.T.cells + cellIdx;
					if (old == 0xFFFFFFFFu) { old = __ldcg(&cell->countMinus1) + 1u; if (old > (uint32_t)g.cap) old = (uint32_t)g.cap; }
					double* lp = A.T.lpos + 3 * (size_t)id;
					lp[0] = cand[0]; lp[1] = cand[1]; lp[2] = cand[2];
					const uint32_t slot = old + rank;
					if (slot < (uint32_t)g.cap) {
						double v[3]; float o[3];
						to_frac(g, cand, v); bead_cell(g, v, o);
						__stcg(&cell->slot[slot], make_float4(o[0], o[1], o[2], __uint_as_float(id)));
						if (rank == 0) {   // one writer per cell: the new count of this round
							uint32_t nc = old + tot + 1u; if (nc > (uint32_t)g.cap) nc = (uint32_t)g.cap;
							cell->countMinus1 = nc - 1u;
						}
					} else {           // cell full: overflow list (rare); the header count is written by the rank-0 bead, if it fits
						atomicExch(&cell->ovfFlagInv, 0u);
						const uint32_t at = atomicAdd(A.T.ovfCount, 1u);
						if (at < A.T.ovfCap) A.T.ovf[at] = id;
					}
					for (int k = 0; k < 3; k++) A.resPos[3 * (size_t)i + k] = cand[k];
				}
			}
			if (lane == 0) {
				A.resTrial[i] = rec.trial;
				myTrials += rec.trial <= T_ ? rec.trial : T_;
				if (conflict) { const uint32_t at = atomicAdd(&A.dirtyCount[round], 1u); A.dirtyList[at] = i; }
			}
			if (conflict)
				for (unsigned c = lane; c < nCta; c += 32) *cluster.map_shared_rank(const_cast<uint32_t*>(&dirtyWord[par]), c) = 1u;
		}
		if (i == 0 && lane == 0) { A.roundAcc[round] = acc; A.roundProg[round] = prog; }
		cluster_arrive(); cluster_wait();
		const long long tk2 = clock64();

		// ---------------- clean-up (rare): the shared serial path through global memory
		if (dirtyWord[par] != 0u) {
			if (i == 0) cleanup_round(A, round, lane);
			cluster_arrive(); cluster_wait();
			cleanRounds++;
			if (active) {
				const uint32_t t = __ldcg(&A.resTrial[i]);
				accepted = (t >= 1u && t <= T_) ? t : 0u;
				if (accepted) for (int k = 0; k < 3; k++) cand[k] = __ldcg(&A.resPos[3 * (size_t)i + k]);
			}
			acc = __ldcg(&A.roundAcc[round]); prog = __ldcg(&A.roundProg[round]);
		}
		// ---------------- commit into the register-resident chain state (sarw.cpp:270-273)
		if (active) {
			if (accepted) {
				const uint32_t id = idRound + i;
				if (lane == 0) A.prev[id - g.idAtomBase] = cs.lastId;
				for (int k = 0; k < 3; k++) { cs.pen[k] = cs.last[k]; cs.last[k] = cand[k]; }
				cs.penId = cs.lastId; cs.lastId = id; cs.length += 1;
				if (lane == 0) A.chains[i] = cs;
			}
			if (lane == 0) A.resTrial[i] = 0;
		}
		nBeads += acc; lastProg = prog != 0u; round++;
		cycA += tk1 - tk0; cycV += tk2 - tk1; cycC += clock64() - tk2;
	}
	if (lane == 0) {
		if (myTrials) atomicAdd(&A.ctl->trials, myTrials);
		if (myQueries) atomicAdd(&A.ctl->queries, myQueries);
	}
	if (i == 0 && lane == 0) { A.ctl->round = round; A.ctl->nBeads = nBeads; A.ctl->lastProgressed = lastProg ? 1u : 0u; }
	if (timer) { A.ctl->cycPropose += cycA; A.ctl->cycVerify += cycV; A.ctl->cycCleanup += cycC; A.ctl->cleanupRounds += cleanRounds; }
	cluster_arrive(); cluster_wait();                    // no CTA may exit while others can still write into its shared memory
}
