// sarw_device.cuh — device-side building blocks of libsarw_b200 (sm_100a).
//
// Everything that decides accept/reject is IEEE-754 double arithmetic in the reference's operation order
// (file:line citations are into /root/reference); the translation unit is compiled with -fmad=false so that no
// multiply-add is contracted. fp32 is used only as a conservative pre-filter whose undecided band falls back
// to the exact test.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sarw {

// ------------------------------------------------------------------ data layout in HBM
// One 32-byte record per cell (one DRAM sector; four x-adjacent cells share a 128-byte line): up to 3 beads inline.
//   id[k]     lookup index of the bead in slot k, ID_NONE = free. A slot is claimed with one atomicCAS on its id word, so there
//             is no count to keep consistent and a removed bead simply leaves a free slot behind.
//   off[k][d] position of bead k inside the cell along d in units of edge/65536 (the fp32 pre-filter works on the centre of the
//             quantisation step; its undecided band covers the half step, everything inside the band is re-evaluated exactly
//             from lpos in the reference's fp64 arithmetic).
//   ovfInv    0xFFFF = no bead of this cell lives in the overflow table (the memset state), anything else = some do (or did).
// A memset to 0xFF yields "empty cell, not overflowed". Beads that do not fit (more than `cap` in one cell: ~1e-4 of the beads
// at melt density with 2 A cells) go to an open-addressing hash table keyed by the cell index; a query probes it only for
// the visited cells that are flagged.
// Round-1 format for comparison: 128 bytes per cell (header + 7 fp32 slots), 0.32 beads per cell on average: 400 B of table
// per bead, one 128-byte line per visited cell of which 32 bytes were useful (ncu: 1.95x wasted DRAM traffic in K5).
struct __align__(32) Cell {
	uint32_t id[3];
	uint16_t off[3][3];
	uint16_t ovfInv;
};
static_assert(sizeof(Cell) == 32, "cell record must be one 32-byte sector");
constexpr int CELL_WORDS = 8;

constexpr uint32_t ID_NONE = 0xFFFFFFFFu;
constexpr uint32_t ID_TOMB = 0xFFFFFFFEu;   // overflow-table entry whose bead was removed again (probing continues past it)
constexpr uint32_t ID_BUSY = 0xFFFFFFFDu;   // overflow-table entry claimed, record not yet published
constexpr int MAX_CAP = 3;

// Per-chain growth state (sarw.h:81-83 lastIndices / penultimateIndices / chainLengths, plus the two tip
// positions so that a round needs one 64-byte read per chain).
struct __align__(16) ChainState {
	double last[3];
	double pen[3];
	uint32_t lastId;
	uint32_t penId;
	uint32_t length;
	uint32_t flags;
};
static_assert(sizeof(ChainState) == 64, "chain state is 64 bytes");

// Immutable per-context parameters, passed to kernels by value.
struct Geom {
	double L[3], Linv[3];     // boxSize, 1./boxSize (PeriodicLookup.h:10-11,30)
	double thresh;            // minDist
	double a[3], ainv[3];     // cell edge L/S and its inverse
	double sOverL[3];         // S as double (cell-space scale for fractional coordinates)
	double reachCells[3];     // conservative query half-width in cell units (< 0.5)
	int S[3];
	int cap;                  // inline capacity in use (<= 3)
	float qs[3];              // edge / 65536: quantisation step of the in-cell offsets
	float lo2, hi2;           // fp32 pre-filter: d2 < lo2 => certainly within thresh, d2 > hi2 => certainly outside
	double regionReach;       // thresh + margin, Angstrom: halo around the cone circle when staging
	double arcShrink;         // blocked arcs are built for thresh - arcShrink: dominates the error of the staged fp32 positions (quantised offsets)
	double dSinPhi, dCosPhi;  // BondLength*sin(phi), BondLength*cos(phi)   (sarw.cpp:165,178-180)
	double bias[3];
	int biased;
	int useTmap;              // the launch carries a CUtensorMap of the cell table (3-D, box 4 x 4 x 4 cells): regions are staged with one tensor copy where possible
	int boundary, graftLoc, growthOrder;
	int maxTrials;
	uint32_t key0, key1;
	uint32_t nChains, nGraft;
	uint32_t idAtomBase;      // lookup index of atom slot 0 (= nObstacles, sarw.cpp:198)
	uint32_t seedIgnoreId;    // nObstacles-1 or ID_NONE: the reference's "ignoreIndex=-1 + nObstacles" quirk (sarw.cpp:198,353)
	int64_t target;
	int64_t nGraftAtoms;      // threshold of sarw.cpp:232
	// Slab decomposition of ONE melt over several GPUs: this context holds the cell layers (z) [winLo, winLo + winLen)
	// (mod S[2]) = its own slab plus a halo on either side; winLen == S[2] means the whole table (every single-GPU context).
	int winLo, winLen;
	int slabRank, slabRanks;  // slabRanks <= 1: no decomposition
	int tmapZ0, tmapZn;       // cell layers the tensor map of the cell table covers: [tmapZ0, tmapZ0 + tmapZn) (the whole table unless this is an inner slab)
	int winStrict;            // growth launches: a probe outside the window is a protocol error and is counted (seeding and obstacles use the hash table there)
};

struct Tables {
	Cell* cells;
	double* lpos;             // Cartesian position by lookup index, 3 doubles each; all-ones bit pattern = never accepted
	uint4* ovf;               // overflow hash table (open addressing, linear probing), entry = {cell index low word, lookup index,
	                          //   off0 | off1 << 16, off2 | cell index high bits << 16}; all ones = never used
	uint32_t* ovfCount;       // entries ever claimed (monotonic)
	uint32_t ovfMask;         // table size - 1 (power of two)
	uint32_t ovfLimit;        // at most this many entries are claimed (half the table): probing always terminates; beyond it the host fails the call
	unsigned* oowCount;       // probes / removals that fell outside the window of a slab context during growth (must stay 0)
	const void* tmapHost;     // HOST pointer to the 128-byte CUtensorMap of the cell table (launch wrappers copy it into the kernel parameters), or nullptr
};

// ------------------------------------------------------------------ Philox4x32-10 (independent of oracle/philox.h)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
		uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
		uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
		c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	return c;
}
__device__ __forceinline__ uint32_t pick(uint4 w, uint32_t i)
{	return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}
// U[0,1) of growth trial t (1-based): counter (chain, round, (t-1)>>2, 0), word (t-1)&3
__device__ __forceinline__ double u_grow(const Geom& g, uint32_t chain, uint32_t round, uint32_t trial)
{	uint4 w = philox4x32_10(make_uint4(chain, round, (trial - 1u) >> 2, 0u), g.key0, g.key1);
	return (double)pick(w, (trial - 1u) & 3u) * (1.0 / 4294967296.0);
}

// ------------------------------------------------------------------ portable sin/cos
// Fixed sequence of double +,-,* (valid for 0 <= x < 2*pi + small): identical bits on CPU and GPU.
__host__ __device__ __forceinline__ void sincos_portable(double x, double& s, double& c)
{
	const double PIO2_HI = 1.57079632673412561417e+00, PIO2_LO = 6.07710050650619224932e-11;
	const double TWO_OVER_PI = 6.36619772367581382433e-01;
	const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
		S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
	const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
		C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
	int k = (int)(x * TWO_OVER_PI + 0.5);
	double kd = (double)k;
	double r = (x - kd * PIO2_HI) - kd * PIO2_LO;
	double z = r * r;
	double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
	double sr = r + (z * r) * (S1 + z * ps);
	double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
	double cr = 1.0 - (0.5 * z - z * pc);
	switch (k & 3) {
	case 0: s = sr; c = cr; break;
	case 1: s = cr; c = 0.0 - sr; break;
	case 2: s = 0.0 - sr; c = 0.0 - cr; break;
	default: s = 0.0 - cr; c = sr; break;
	}
}

// ------------------------------------------------------------------ geometry in the reference's operation order
__device__ __forceinline__ void normalize3(double v[3])   // v * (1.0 / sqrt(v0*v0 + v1*v1 + v2*v2))
{	double inv = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	v[0] = v[0] * inv; v[1] = v[1] * inv; v[2] = v[2] * inv;
}

struct Frame { double X[3], Y[3], Z[3]; };

// local frame of the cone, sarw.cpp:168-173
__device__ __forceinline__ void cone_frame(const double last[3], const double pen[3], Frame& f)
{
	f.Z[0] = last[0] - pen[0]; f.Z[1] = last[1] - pen[1]; f.Z[2] = last[2] - pen[2];
	normalize3(f.Z);
	if (abs((int)f.Z[0]) > 0) {     // sarw.cpp:170: abs() there is int abs(int)
		f.X[0] = 1.0 - f.Z[0] * f.Z[0]; f.X[1] = 0.0 - f.Z[1] * f.Z[0]; f.X[2] = 0.0 - f.Z[2] * f.Z[0];
	} else {                        // sarw.cpp:171
		f.X[0] = 0.0 - f.Z[0] * f.Z[1]; f.X[1] = 1.0 - f.Z[1] * f.Z[1]; f.X[2] = 0.0 - f.Z[2] * f.Z[1];
	}
	normalize3(f.X);
	f.Y[0] = f.Z[1] * f.X[2] - f.Z[2] * f.X[1];
	f.Y[1] = f.Z[2] * f.X[0] - f.Z[0] * f.X[2];
	f.Y[2] = f.Z[0] * f.X[1] - f.Z[1] * f.X[0];
	normalize3(f.Y);
}

// candidate of sarw.cpp:177-184 for u01 = U[0,1)
__device__ __forceinline__ void cone_candidate(const Geom& g, const Frame& f, const double last[3], double u01, double out[3])
{
	double theta = (0.0 + (359.0 - 0.0) * u01) * 3.14159265358979323846 / 180;
	double st, ct;
	sincos_portable(theta, st, ct);
	double a = g.dSinPhi * ct, b = g.dSinPhi * st, c = g.dCosPhi;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double np = (f.X[k] * a + f.Y[k] * b) + f.Z[k] * c;
		out[k] = np + last[k];
	}
}

// fractional coordinates of PeriodicLookup.h:38-41 / :53-55
__device__ __forceinline__ void to_frac(const Geom& g, const double pos[3], double v[3])
{
#pragma unroll
	for (int k = 0; k < 3; k++) { double t = pos[k] * g.Linv[k]; v[k] = t - floor(t); }
}

// floor() without the conversion unit. On B200 FRND/F2I/I2F/F2F on doubles issue on the XU pipe (16 lanes/clk/SM); the overlap-query
// kernel was bound by it (ncu: xu 92 %). Adding and subtracting 1.5*2^52 rounds to the nearest integer with two DADDs (exact, |t| < 2^51),
// one compare turns that into floor; the integer itself is the low word of the biased sum. Same values as floor(), bit for bit.
constexpr double FLOOR_MAGIC = 6755399441055744.0;
__device__ __forceinline__ double floor_noxu(double t)
{
	if (!(fabs(t) < 2251799813685248.0)) return floor(t);   // out of range of the trick (or NaN): the plain instruction
	const double r = (t + FLOOR_MAGIC) - FLOOR_MAGIC;
	return r > t ? r - 1.0 : r;
}
// floor(t) as int and as double, for |t| < 2^31
__device__ __forceinline__ int floor_int_noxu(double t, double& fl)
{
	const double tm = t + FLOOR_MAGIC;
	const double r = tm - FLOOR_MAGIC;
	const int i = __double2loint(tm);
	const bool down = r > t;
	fl = down ? r - 1.0 : r;
	return down ? i - 1 : i;
}
__device__ __forceinline__ void to_frac_noxu(const Geom& g, const double pos[3], double v[3])
{
#pragma unroll
	for (int k = 0; k < 3; k++) { double t = pos[k] * g.Linv[k]; v[k] = t - floor_noxu(t); }
}

// the per-pair decision of PeriodicLookup.h:67-71 between candidate (fractional v) and stored bead (Cartesian p)
__device__ __forceinline__ bool exact_pair(const Geom& g, const double v[3], const double* __restrict__ p)
{
	double c[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double t = __ldcg(p + k) * g.Linv[k];
		double vp = t - floor(t);
		double dv = v[k] - vp;
		dv = dv - floor(0.5 + dv);
		c[k] = g.L[k] * dv;
	}
	double dist = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
	return dist <= g.thresh;
}

// same decision with the stored bead's Cartesian position already in registers
__device__ __forceinline__ bool exact_pair_reg(const Geom& g, const double v[3], const double p[3])
{
	double c[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double t = p[k] * g.Linv[k];
		double vp = t - floor(t);
		double dv = v[k] - vp;
		dv = dv - floor(0.5 + dv);
		c[k] = g.L[k] * dv;
	}
	double dist = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
	return dist <= g.thresh;
}

// fixed walls (sarw.cpp:201-203), zCenter half-space rule (:206-211) and the growthBias extension
__device__ __forceinline__ bool walls_reject(const Geom& g, const double pos[3])
{
	if (g.boundary == 1) return pos[2] < 0 || pos[2] > g.L[2];
	if (g.boundary == 2) return pos[1] < 0 || pos[1] > g.L[1];
	if (g.boundary == 3) return pos[0] < 0 || pos[0] > g.L[0];
	return false;
}
__device__ __forceinline__ bool growth_rules_reject(const Geom& g, const double pos[3], const double last[3])
{
	if (g.biased) {
		double sx = pos[0] - last[0], sy = pos[1] - last[1], sz = pos[2] - last[2];
		if ((g.bias[0] * sx + g.bias[1] * sy) + g.bias[2] * sz < 0) return true;
	}
	if (walls_reject(g, pos)) return true;
	if (g.graftLoc == 2) {
		double lastPos = last[2] - 0.5 * g.L[2];
		double newPos = pos[2] - 0.5 * g.L[2];
		if (lastPos * newPos < 0) return true;
	}
	return false;
}

// ------------------------------------------------------------------ cell addressing
__device__ __forceinline__ int wrap_cell(int c, int S)
{	// callers are at most one period off: the integer division (I2F + MUFU.RCP + F2I on the XU pipe, ~70 cycles) is the rare path
	if (c < 0) c += S; else if (c >= S) c -= S;
	if ((unsigned)c >= (unsigned)S) { c %= S; if (c < 0) c += S; }
	return c;
}
__device__ __forceinline__ int64_t cell_index(const Geom& g, int wx, int wy, int wz)
{	return wx + (int64_t)g.S[0] * (wy + (int64_t)g.S[1] * wz);
}

// ------------------------------------------------------------------ slab decomposition: which cell layers this context holds, who owns a layer
__device__ __forceinline__ bool windowed(const Geom& g) { return g.winLen < g.S[2]; }
__device__ __forceinline__ bool in_window(const Geom& g, int wz)
{	int d = wz - g.winLo;
	d += d < 0 ? g.S[2] : 0;
	return d < g.winLen;
}
__host__ __device__ __forceinline__ int slab_lo(int S2, int ranks, int r) { return (int)(((long long)r * S2) / ranks); }   // slab r = layers [slab_lo(r), slab_lo(r + 1))
__host__ __device__ __forceinline__ int slab_owner_of(int S2, int ranks, int wz)
{	int r = (int)(((long long)wz * ranks) / S2);
	while (r > 0 && wz < slab_lo(S2, ranks, r)) --r;
	while (r + 1 < ranks && wz >= slab_lo(S2, ranks, r + 1)) ++r;
	return r;
}
__device__ __forceinline__ int slab_owner(const Geom& g, int wz) { return slab_owner_of(g.S[2], g.slabRanks, wz); }
// window of rank r: its slab plus `halo` layers on either side, as (first layer, number of layers), first + count may exceed S2 (wraps);
// the whole table when the window would cover it anyway
__host__ __device__ __forceinline__ void slab_window(int S2, int ranks, int r, int halo, int& winLo, int& winLen)
{	const int lo = slab_lo(S2, ranks, r), hi = slab_lo(S2, ranks, r + 1);
	if (hi - lo + 2 * halo < S2) { winLo = ((lo - halo) % S2 + S2) % S2; winLen = hi - lo + 2 * halo; }
	else { winLo = 0; winLen = S2; }
}

// cell + quantised in-cell offset of a bead from its fractional coordinates
__device__ __forceinline__ int64_t bead_cell(const Geom& g, const double v[3], uint32_t q16[3], int c[3])
{
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double s = v[k] * g.sOverL[k];
		int ci = (int)floor(s);
		if (ci >= g.S[k]) ci = g.S[k] - 1;
		if (ci < 0) ci = 0;
		c[k] = ci;
		int q = (int)((s - (double)ci) * 65536.0);
		q16[k] = (uint32_t)(q < 0 ? 0 : (q > 65535 ? 65535 : q));
	}
	return cell_index(g, c[0], c[1], c[2]);
}
__device__ __forceinline__ int64_t bead_cell(const Geom& g, const double v[3], uint32_t q16[3])
{	int c[3];
	return bead_cell(g, v, q16, c);
}

// the whole 32-byte record with one 256-bit load (SASS LDG.E.256), straight from L2
__device__ __forceinline__ void ld_cell(const Cell* cell, uint32_t (&w)[CELL_WORDS])
{
	asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(cell));
}
// quantised offset of slot k along d from the record's words (k, d compile-time constants after unrolling)
__device__ __forceinline__ uint32_t cell_off(const uint32_t (&w)[CELL_WORDS], int k, int d)
{	const int e = 3 * k + d;
	return (w[3 + (e >> 1)] >> (16 * (e & 1))) & 0xFFFFu;
}
__device__ __forceinline__ bool cell_flagged(const uint32_t (&w)[CELL_WORDS]) { return (w[7] >> 16) != 0xFFFFu; }
// centre of the quantisation step, in Angstrom from the cell's low corner: (q + 0.5) * step. The integer -> float conversion (I2F)
// issues on the XU pipe (16 lanes/clk/SM; 72 of them per candidate made K5 slower with the compact cells than with fp32 offsets),
// so q (< 2^16) is placed in the mantissa of 2^23 instead: 2^23 + q - (2^23 - 0.5) = q + 0.5 exactly, one LOP3 and one FADD.
__device__ __forceinline__ float dec_off(const Geom& g, uint32_t q, int d) { return (__uint_as_float(0x4B000000u | q) - 8388607.5f) * g.qs[d]; }

// ------------------------------------------------------------------ overflow table (beads whose cell was full)
__device__ __forceinline__ uint32_t ovf_hash(int64_t cidx)
{	unsigned long long x = (unsigned long long)cidx * 0x9E3779B97F4A7C15ull;
	return (uint32_t)(x >> 32);
}
// visit every published entry of cell `cidx`: fn(id, q0, q1, q2) returns true to stop early
template <typename F>
__device__ __forceinline__ void ovf_for_cell(const Tables& T, int64_t cidx, F&& fn)
{
	const uint32_t lo = (uint32_t)cidx, hi = (uint32_t)((unsigned long long)cidx >> 32);
	uint32_t h = ovf_hash(cidx) & T.ovfMask;
	for (uint32_t n = 0; n <= T.ovfMask; ++n, h = (h + 1u) & T.ovfMask) {
		const uint4 e = __ldcg(T.ovf + h);
		if (e.y == ID_NONE) return;                 // never used: end of every probe sequence through here
		if (e.y >= ID_BUSY) continue;               // being written (a bead of the running round: nobody needs it yet) or removed
		if (e.x == lo && (e.w >> 16) == hi && fn(e.y, e.z & 0xFFFFu, e.z >> 16, e.w & 0xFFFFu)) return;
	}
}
__device__ inline void ovf_insert(const Tables& T, int64_t cidx, uint32_t id, const uint32_t q16[3])
{
	if (atomicAdd(T.ovfCount, 1u) >= T.ovfLimit) return;   // the host sees ovfCount > ovfLimit and fails the call
	uint32_t h = ovf_hash(cidx) & T.ovfMask;
	while (atomicCAS(&T.ovf[h].y, ID_NONE, ID_BUSY) != ID_NONE) h = (h + 1u) & T.ovfMask;
	// published with ONE 16-byte store: a concurrent reader sees either the busy entry or the complete one
	__stcg(T.ovf + h, make_uint4((uint32_t)cidx, id, q16[0] | (q16[1] << 16), q16[2] | ((uint32_t)((unsigned long long)cidx >> 32) << 16)));
}
__device__ inline void ovf_remove(const Tables& T, int64_t cidx, uint32_t id)
{
	uint32_t h = ovf_hash(cidx) & T.ovfMask;
	for (uint32_t n = 0; n <= T.ovfMask; ++n, h = (h + 1u) & T.ovfMask) {
		const uint32_t y = __ldcg(&T.ovf[h].y);
		if (y == ID_NONE) return;
		if (y == id) { atomicExch(&T.ovf[h].y, ID_TOMB); return; }   // tombstone: entries are never moved, so concurrent probes stay valid
	}
}

// ------------------------------------------------------------------ overlap query against the cell list
// fp32 pre-filter + exact fall-back for one stored bead; (dx, dy, dz) = bead - candidate in fp32
__device__ __forceinline__ bool pair_hits(const Geom& g, const Tables& T, const double v[3], float dx, float dy, float dz, uint32_t id)
{
	const float d2 = dx * dx + dy * dy + dz * dz;
	if (d2 < g.lo2) return true;
	return d2 <= g.hi2 && exact_pair(g, v, T.lpos + 3 * (size_t)id);
}

// Does any stored bead with lookup index in [idLo, idHi) and != ignoreId lie within thresh of the candidate?
// GROUP lanes share the (<= 8) cells of the query: lane `sub` takes cells sub, sub+GROUP, ... ; the caller ORs
// the per-lane results.
template <int GROUP>
__device__ __forceinline__ bool probe_cells(const Geom& g, const Tables& T, const double v[3],
	uint32_t idLo, uint32_t idHi, uint32_t ignoreId, int sub)
{
	int c0[3], n[3];
	float q[3][2];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double s = v[k] * g.sOverL[k];
		int lo = (int)floor(s - g.reachCells[k]);
		int hi = (int)floor(s + g.reachCells[k]);
		c0[k] = lo; n[k] = hi - lo + 1;   // 1 or 2 (reachCells < 0.5)
		q[k][0] = (float)(((double)lo - s) * g.a[k]);
		q[k][1] = (float)(((double)(lo + 1) - s) * g.a[k]);
	}
	bool hit = false;
	const int total = n[0] * n[1] * n[2];
	for (int ci = sub; ci < total && !hit; ci += GROUP) {
		const int ix = ci % n[0], rest = ci / n[0];
		const int iy = rest % n[1], iz = rest / n[1];
		const int wz = wrap_cell(c0[2] + iz, g.S[2]);
		const int64_t cidx = cell_index(g, wrap_cell(c0[0] + ix, g.S[0]), wrap_cell(c0[1] + iy, g.S[1]), wz);
		const float qx = q[0][ix], qy = q[1][iy], qz = q[2][iz];
		if (windowed(g) && !in_window(g, wz)) {   // slab context, layer held by another rank: only seeds / obstacles are known here (hash table)
			if (g.winStrict) atomicAdd(T.oowCount, 1u);
			ovf_for_cell(T, cidx, [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
				if (id < idLo || id >= idHi || id == ignoreId) return false;
				hit = pair_hits(g, T, v, qx + dec_off(g, o0, 0), qy + dec_off(g, o1, 1), qz + dec_off(g, o2, 2), id);
				return hit;
			});
			continue;
		}
		uint32_t w[CELL_WORDS];
		ld_cell(T.cells + cidx, w);
#pragma unroll
		for (int s = 0; s < MAX_CAP; ++s) {
			const uint32_t id = w[s];
			if (hit || id < idLo || id >= idHi || id == ignoreId) continue;   // idHi <= 0xFFFFFFF0: free slots never pass
			hit = pair_hits(g, T, v, qx + dec_off(g, cell_off(w, s, 0), 0), qy + dec_off(g, cell_off(w, s, 1), 1), qz + dec_off(g, cell_off(w, s, 2), 2), id);
		}
		if (!hit && cell_flagged(w))
			ovf_for_cell(T, cidx, [&](uint32_t id, uint32_t o0, uint32_t o1, uint32_t o2) {
				if (id < idLo || id >= idHi || id == ignoreId) return false;
				hit = pair_hits(g, T, v, qx + dec_off(g, o0, 0), qy + dec_off(g, o1, 1), qz + dec_off(g, o2, 2), id);
				return hit;
			});
	}
	return hit;
}

// Cells are written through the generic proxy (the inserts below) and read in later rounds by bulk / tensor copies, i.e. through the
// async proxy; the PTX memory model orders the two only across a proxy fence. Every thread that wrote a cell issues it right after
// its stores; the round's barrier (release / acquire) then carries the ordering to the thread that issues the copy. (The LDGSTS
// staging variant reads through the generic proxy and needs none.)
__device__ __forceinline__ void fence_cells_for_tma()
{
#ifndef SARW_NO_PROXY_FENCE
	asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
}

__device__ __forceinline__ void st_off16(Cell* cell, int k, const uint32_t q16[3])
{
	unsigned short* o = reinterpret_cast<unsigned short*>(&cell->off[k][0]);
	o[0] = (unsigned short)q16[0]; o[1] = (unsigned short)q16[1]; o[2] = (unsigned short)q16[2];
}

// insert a bead (PeriodicLookup.h:37-47 + the atom store of sarw.h:123-126); any number of threads may insert concurrently.
// The slot is claimed by a compare-and-swap on its id word; the offsets follow with plain stores: a bead inserted in the running
// round is never read before the next grid-wide barrier (readers filter on id < first id of the round).
__device__ __forceinline__ void insert_bead(const Geom& g, const Tables& T, const double pos[3], uint32_t id)
{
	double v[3]; uint32_t q16[3];
	to_frac(g, pos, v);
	const int64_t ci = bead_cell(g, v, q16);
	double* lp = T.lpos + 3 * (size_t)id;
	lp[0] = pos[0]; lp[1] = pos[1]; lp[2] = pos[2];
	if (windowed(g) && !in_window(g, (int)(ci / ((int64_t)g.S[0] * g.S[1])))) {
		// slab context, layer held by another rank: growth beads are only recorded in lpos (replicated); seeds and obstacles, which every
		// rank places and checks on its own, are kept in the hash table
		if (!g.winStrict) ovf_insert(T, ci, id, q16);
		return;
	}
	Cell* cell = T.cells + ci;
	for (int k = 0; k < g.cap; ++k)
		if (atomicCAS(&cell->id[k], ID_NONE, id) == ID_NONE) { st_off16(cell, k, q16); fence_cells_for_tma(); return; }
	cell->ovfInv = 0;              // idempotent plain store: every writer stores the same value
	ovf_insert(T, ci, id, q16);
	fence_cells_for_tma();
}

// remove a bead again (conflict clean-up; several warps may do this at once for beads of different chains)
__device__ inline void remove_bead(const Geom& g, const Tables& T, uint32_t id)
{
	double* lp = T.lpos + 3 * (size_t)id;
	double pos[3] = { __ldcg(lp), __ldcg(lp + 1), __ldcg(lp + 2) };
	double v[3]; uint32_t q16[3];
	to_frac(g, pos, v);
	const int64_t ci = bead_cell(g, v, q16);
	if (windowed(g) && !in_window(g, (int)(ci / ((int64_t)g.S[0] * g.S[1])))) {
		if (!g.winStrict) ovf_remove(T, ci, id);
	} else {
		Cell* cell = T.cells + ci;
		bool found = false;
		for (int k = 0; k < MAX_CAP && !found; ++k)
			if (__ldcg(&cell->id[k]) == id) { atomicExch(&cell->id[k], ID_NONE); found = true; }   // the slot is free for the next insert
		if (!found) ovf_remove(T, ci, id);
	}
	unsigned long long ones = ~0ull;
	lp[0] = __longlong_as_double((long long)ones); lp[1] = lp[0]; lp[2] = lp[0];
	__threadfence();
	fence_cells_for_tma();
}

// ------------------------------------------------------------------ grid-wide barrier for the persistent kernel
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{	unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// One arrival counter that only ever grows (the launch wrapper zeroes it): barrier k of a launch is passed when the counter reaches
// k * nBlocks. Per CTA one release-reduction and an acquire spin; the CTA barrier before it makes the release cover the stores of the whole
// CTA (cumulativity), the one after it hands the acquired view to the other threads. The first version (arrival counter reset by the
// last CTA + a generation word, three fences) cost ~5k cycles per barrier: a third of a C3 round with its 3-6 barriers.
__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned& target, unsigned nBlocks)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		target += nBlocks;
		asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(count) : "memory");
		while ((int)(ld_acquire_u32(count) - target) < 0) { }
	}
	__syncthreads();
}

} // namespace sarw
