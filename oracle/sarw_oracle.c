/* sarw_oracle.c — see sarw_oracle.h. TEST INFRASTRUCTURE ONLY; compiled with -ffp-contract=off, no -march.
 * All file:line citations are into /root/reference. */
#include "sarw_oracle.h"
#include "philox.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define AVOGADRO 6.022e23      /* sarw.h:4 */
#define BOND_LENGTH 1.54       /* sarw.h:5 */
#define ATOM_MASS 14.0         /* sarw.h:6 */

/* ------------------------------------------------------------------ portable sin/cos
 * A fixed sequence of IEEE-754 double +,-,* (no FMA, no libm) so that CPU and GPU agree bit for bit.
 * Valid for 0 <= x < 2*pi + small (all this path needs: theta in [0, 359 deg), phi = 70.5 deg).
 * Reduction by quadrant with a two-term pi/2, then the classic degree-13 / degree-14 minimax kernels
 * (coefficients as published in Sun's fdlibm k_sin.c / k_cos.c). Max observed error < 2 ulp vs libm. */
static const double PIO2_HI = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
static const double PIO2_LO = 6.07710050650619224932e-11; /* pi/2 - PIO2_HI */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
	S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
	C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

void oracle_sincos_portable(double x, double* s, double* c)
{
	int k = (int)(x * TWO_OVER_PI + 0.5);
	double kd = (double)k;
	double r = (x - kd * PIO2_HI) - kd * PIO2_LO;
	double z = r * r;
	double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
	double sr = r + (z * r) * (S1 + z * ps);
	double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
	double cr = 1.0 - (0.5 * z - z * pc);
	switch (k & 3) {
	case 0: *s = sr; *c = cr; break;
	case 1: *s = cr; *c = 0.0 - sr; break;
	case 2: *s = 0.0 - sr; *c = 0.0 - cr; break;
	default: *s = 0.0 - cr; *c = sr; break;
	}
}

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { oracle_philox4x32_10(ctr, key, out); }

/* ------------------------------------------------------------------ PeriodicLookup (PeriodicLookup.h) */
struct oracle_lookup {
	double box[3], boxInv[3], thresh;   /* PeriodicLookup.h:10-11 */
	int S[3];                           /* :12 */
	double* pts;                        /* :13 points (fractional coordinates) */
	int64_t nPts, capPts;
	int32_t* head;                      /* :14 indices, as per-cell singly linked lists */
	int32_t* next;
};

oracle_lookup* oracle_lookup_create(double lx, double ly, double lz, double thresh)
{
	oracle_lookup* lk = (oracle_lookup*)calloc(1, sizeof *lk);
	lk->box[0] = lx; lk->box[1] = ly; lk->box[2] = lz; lk->thresh = thresh;
	for (int k = 0; k < 3; k++) {                 /* :28-31 */
		lk->S[k] = (int)floor(lk->box[k] / thresh);
		lk->boxInv[k] = 1. / lk->box[k];
	}
	size_t nCells = (size_t)lk->S[0] * lk->S[1] * lk->S[2];   /* :33 */
	lk->head = (int32_t*)malloc(nCells * sizeof(int32_t));
	memset(lk->head, 0xff, nCells * sizeof(int32_t));
	return lk;
}
void oracle_lookup_destroy(oracle_lookup* lk) { if (lk) { free(lk->pts); free(lk->head); free(lk->next); free(lk); } }

static size_t mesh_index(const oracle_lookup* lk, int iv[3])   /* :16-22, single wrap */
{
	for (int k = 0; k < 3; k++) {
		if (iv[k] < 0) iv[k] += lk->S[k];
		if (iv[k] >= lk->S[k]) iv[k] -= lk->S[k];
	}
	return iv[0] + lk->S[0] * (size_t)(iv[1] + lk->S[1] * iv[2]);
}

static void lookup_add_point(oracle_lookup* lk, const double pos[3])   /* :37-47 */
{
	double v[3]; int iv[3];
	for (int k = 0; k < 3; k++) {
		v[k] = pos[k] * lk->boxInv[k];
		v[k] -= floor(v[k]);
		iv[k] = (int)floor(v[k] * lk->S[k] + 0.5);
	}
	if (lk->nPts == lk->capPts) {
		lk->capPts = lk->capPts ? 2 * lk->capPts : 1024;
		lk->pts = (double*)realloc(lk->pts, (size_t)lk->capPts * 3 * sizeof(double));
		lk->next = (int32_t*)realloc(lk->next, (size_t)lk->capPts * sizeof(int32_t));
	}
	size_t cell = mesh_index(lk, iv);
	lk->next[lk->nPts] = lk->head[cell];
	lk->head[cell] = (int32_t)lk->nPts;
	memcpy(lk->pts + 3 * lk->nPts, v, sizeof v);
	lk->nPts++;
}

/* the per-pair test of :67-71 on fractional coordinates */
static int pair_within(const oracle_lookup* lk, const double v[3], const double* p)
{
	double dv[3];
	for (int k = 0; k < 3; k++) { dv[k] = v[k] - p[k]; dv[k] -= floor(0.5 + dv[k]); }
	double a = lk->box[0] * dv[0], b = lk->box[1] * dv[1], c = lk->box[2] * dv[2];
	double dist = sqrt(a * a + b * b + c * c);
	return dist <= lk->thresh;
}

static int lookup_find(const oracle_lookup* lk, const double pos[3], int ignoreIndex)   /* :51-75 */
{
	int ivMin[3], ivMax[3]; double v[3];
	for (int k = 0; k < 3; k++) {
		v[k] = pos[k] * lk->boxInv[k];
		v[k] -= floor(v[k]);
		double thresh_k = lk->thresh * lk->boxInv[k];
		ivMin[k] = (int)floor((v[k] - thresh_k) * lk->S[k] + 0.5);
		ivMax[k] = (int)floor((v[k] + thresh_k) * lk->S[k] + 0.5);
	}
	for (int i0 = ivMin[0]; i0 <= ivMax[0]; i0++)
	for (int i1 = ivMin[1]; i1 <= ivMax[1]; i1++)
	for (int i2 = ivMin[2]; i2 <= ivMax[2]; i2++) {
		int iv[3] = { i0, i1, i2 };
		for (int32_t idx = lk->head[mesh_index(lk, iv)]; idx >= 0; idx = lk->next[idx])
			if (ignoreIndex < 0 || idx != ignoreIndex)
				if (pair_within(lk, v, lk->pts + 3 * (size_t)idx)) return 1;
	}
	return 0;
}

void oracle_lookup_add(oracle_lookup* lk, const double* xyz, int64_t n)
{	for (int64_t i = 0; i < n; i++) lookup_add_point(lk, xyz + 3 * i);
}
void oracle_lookup_find(const oracle_lookup* lk, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit)
{	for (int64_t i = 0; i < n; i++) hit[i] = (uint8_t)lookup_find(lk, xyz + 3 * i, ignore ? (int)ignore[i] : -1);
}
void oracle_lookup_find_bruteforce(const oracle_lookup* lk, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit)
{
	for (int64_t i = 0; i < n; i++) {
		double v[3];
		for (int k = 0; k < 3; k++) { v[k] = xyz[3 * i + k] * lk->boxInv[k]; v[k] -= floor(v[k]); }
		int ig = ignore ? (int)ignore[i] : -1, h = 0;
		for (int64_t j = 0; j < lk->nPts && !h; j++)
			if (ig < 0 || (int)j != ig) h = pair_within(lk, v, lk->pts + 3 * j);
		hit[i] = (uint8_t)h;
	}
}

/* ------------------------------------------------------------------ SARW state (sarw.h:59-94) */
struct oracle_walk {
	oracle_params p;
	oracle_lookup* plook;
	int nObstacles;
	const double* grafts; int64_t nGrafts;
	double* pos; int32_t* chainID; int32_t* type; int32_t* accTrial; int64_t nAtoms, capAtoms;   /* polymerChains */
	int64_t* bonds; int64_t nBonds, capBonds;                                                   /* listBonds */
	int32_t* lastIdx; int32_t* penIdx; int32_t* chainLen;                                        /* sarw.h:81-83 */
	int64_t* countAfterRound; int32_t rounds, capRounds;
	int initiated, lastProgressed;
	uint64_t trials, seedTrials;
	double seconds;
	double sinPhi, cosPhi;
};

int64_t oracle_target_count(const double box[3], double targetMassDensity)   /* sarw.h:103-104 */
{
	int64_t vol = (int64_t)(box[0] * box[1] * box[2]);
	return (int64_t)nearbyint((vol * targetMassDensity * AVOGADRO * 1e-24) / ATOM_MASS);
}

static void add_atom(oracle_walk* w, int type, int chainID, const double pos[3], int trial)   /* sarw.h:123-126 */
{
	if (w->nAtoms == w->capAtoms) {
		w->capAtoms = w->capAtoms ? 2 * w->capAtoms : 4096;
		w->pos = (double*)realloc(w->pos, (size_t)w->capAtoms * 3 * sizeof(double));
		w->chainID = (int32_t*)realloc(w->chainID, (size_t)w->capAtoms * sizeof(int32_t));
		w->type = (int32_t*)realloc(w->type, (size_t)w->capAtoms * sizeof(int32_t));
		w->accTrial = (int32_t*)realloc(w->accTrial, (size_t)w->capAtoms * sizeof(int32_t));
	}
	memcpy(w->pos + 3 * w->nAtoms, pos, 3 * sizeof(double));
	w->chainID[w->nAtoms] = chainID; w->type[w->nAtoms] = type; w->accTrial[w->nAtoms] = trial;
	w->nAtoms++;
	lookup_add_point(w->plook, pos);
}
static void add_bond(oracle_walk* w, int64_t a, int64_t b)
{
	if (w->nBonds == w->capBonds) {
		w->capBonds = w->capBonds ? 2 * w->capBonds : 4096;
		w->bonds = (int64_t*)realloc(w->bonds, (size_t)w->capBonds * 2 * sizeof(int64_t));
	}
	w->bonds[2 * w->nBonds] = a; w->bonds[2 * w->nBonds + 1] = b; w->nBonds++;
}

static void trig_sincos(int trig, double x, double* s, double* c)
{
	if (trig == ORACLE_TRIG_PORTABLE) oracle_sincos_portable(x, s, c);
	else { *s = sin(x); *c = cos(x); }
}

static void normalize3(double v[3])   /* stand-in vector3.h: v * (1.0 / sqrt(v0*v0 + v1*v1 + v2*v2)) */
{
	double inv = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	v[0] = v[0] * inv; v[1] = v[1] * inv; v[2] = v[2] * inv;
}

/* sarw.cpp:159-186 with theta = uniform(0,359)*M_PI/180 from u01 (sarw.cpp:177) */
static void cone_pos(const double last[3], const double pen[3], double u01, int trig, double sinPhi, double cosPhi, double out[3])
{
	double Z[3] = { last[0] - pen[0], last[1] - pen[1], last[2] - pen[2] }, X[3], Y[3];   /* :168 */
	normalize3(Z);
	if (abs((int)Z[0]) > 0) {           /* :170 — unqualified abs() binds to int abs(int) */
		X[0] = 1.0 - Z[0] * Z[0]; X[1] = 0.0 - Z[1] * Z[0]; X[2] = 0.0 - Z[2] * Z[0];
	} else {                            /* :171 */
		X[0] = 0.0 - Z[0] * Z[1]; X[1] = 1.0 - Z[1] * Z[1]; X[2] = 0.0 - Z[2] * Z[1];
	}
	normalize3(X);                      /* :172 */
	Y[0] = Z[1] * X[2] - Z[2] * X[1]; Y[1] = Z[2] * X[0] - Z[0] * X[2]; Y[2] = Z[0] * X[1] - Z[1] * X[0];
	normalize3(Y);                      /* :173 */
	double theta = (0.0 + (359.0 - 0.0) * u01) * M_PI / 180;   /* :177 */
	double st, ct;
	trig_sincos(trig, theta, &st, &ct);
	double distance = BOND_LENGTH;
	double a = distance * sinPhi * ct, b = distance * sinPhi * st, c = distance * cosPhi;   /* :178-180 */
	for (int k = 0; k < 3; k++) {
		double np = (X[k] * a + Y[k] * b) + Z[k] * c;
		out[k] = np + last[k];          /* :184 */
	}
}

void oracle_cone_pos(const double last[3], const double pen[3], double u01, int trig, double out[3])
{
	double phi = (180 - 109.5) * M_PI / 180, sp, cp;
	trig_sincos(trig, phi, &sp, &cp);
	cone_pos(last, pen, u01, trig, sp, cp, out);
}

/* sarw.cpp:193-214 for one atom; `step` is used only by the growthBias extension */
static int check_collision(const oracle_walk* w, const double pos[3], int ignoreIndex)
{
	const oracle_params* p = &w->p;
	if (lookup_find(w->plook, pos, ignoreIndex + w->nObstacles)) return 1;           /* :198 */
	if (p->boundary == ORACLE_BOUNDARY_PPF && (pos[2] < 0 || pos[2] > p->box[2])) return 1;   /* :201 */
	if (p->boundary == ORACLE_BOUNDARY_PFP && (pos[1] < 0 || pos[1] > p->box[1])) return 1;   /* :202 */
	if (p->boundary == ORACLE_BOUNDARY_FPP && (pos[0] < 0 || pos[0] > p->box[0])) return 1;   /* :203 */
	if (p->graftLoc == ORACLE_GRAFT_ZCENTER && ignoreIndex > -1) {                   /* :206-211 */
		double lastPos = w->pos[3 * (size_t)ignoreIndex + 2] - 0.5 * p->box[2];
		double newPos = pos[2] - 0.5 * p->box[2];
		if (lastPos * newPos < 0) return 1;
	}
	return 0;
}

/* sarw.cpp:329-357 */
static int random_seed(oracle_walk* w, int iChain, double p0[3], double p1[3], int* trialOut)
{
	const oracle_params* p = &w->p;
	int iTrial = 0;
	while (iTrial++ < p->maxTrials) {
		double u[6];
		oracle_u_seed(p->seed, (uint32_t)iChain, (uint32_t)iTrial, u);
		w->seedTrials++;
		/* :336-339 — the three draws are constructor arguments; g++ evaluates them right to left, so the
		 * first draw lands in z (observed on oracle/_ref; documented in DESIGN.md "RNG") */
		for (int k = 0; k < 3; k++) p0[k] = 0.0 + (p->box[k] - 0.0) * u[2 - k];
		if (p->graftLoc == ORACLE_GRAFT_FILE && iChain < p->nGraftChains)             /* :342-343 */
			for (int k = 0; k < 3; k++) p0[k] = w->grafts[3 * (size_t)iChain + k];
		else if (p->graftLoc == ORACLE_GRAFT_Z && iChain < p->nGraftChains && p->boundary == ORACLE_BOUNDARY_PPF)
			p0[2] = iChain % 2 ? 0 : p->box[2];                                       /* :344-345 */
		else if (p->graftLoc == ORACLE_GRAFT_ZCENTER && iChain < p->nGraftChains && p->boundary == ORACLE_BOUNDARY_PPF)
			p0[2] = 0.5 * p->box[2];                                                  /* :346-347 */
		double step[3] = { -.5 + (.5 - -.5) * u[5], -.5 + (.5 - -.5) * u[4], -.5 + (.5 - -.5) * u[3] };   /* :150, same right-to-left order */
		normalize3(step);                                                             /* :151 */
		for (int k = 0; k < 3; k++) p1[k] = p0[k] + step[k] * BOND_LENGTH;             /* :350,352 */
		if (!check_collision(w, p0, -1) && !check_collision(w, p1, -1)) break;        /* :353 */
	}
	*trialOut = iTrial;
	if (iTrial >= p->maxTrials) return 0;                                             /* :355 */
	return 1;
}

/* sarw.cpp:290-319 */
static int initiate_chain(oracle_walk* w)
{
	for (int iChain = 0; iChain < w->p.nChains; ++iChain) {
		double p0[3], p1[3]; int trial;
		if (random_seed(w, iChain, p0, p1, &trial)) {
			add_atom(w, 1, iChain + 1, p0, trial);                                    /* :301-304 */
			add_atom(w, 2, iChain + 1, p1, trial);
			w->lastIdx[iChain] = (int32_t)w->nAtoms - 1;                              /* :306-307 */
			w->penIdx[iChain] = (int32_t)w->nAtoms - 2;
			add_bond(w, w->penIdx[iChain] + 1, w->lastIdx[iChain] + 1);               /* :309 */
			w->chainLen[iChain] += 2;
		} else return 0;                                                              /* :313-314 */
	}
	return 1;
}

/* sarw.cpp:220-283 */
static int propagate_chain(oracle_walk* w, uint32_t round)
{
	const oracle_params* p = &w->p;
	int startChain, endChain, flag = 1;
	int64_t target = oracle_target_count(p->box, p->targetMassDensity);
	if (p->growthOrder == ORACLE_ORDER_GRAFTED && p->nGraftChains > 0) {              /* :230-244 */
		int64_t nGraftAtoms = (p->nGraftChains * target / p->nChains) + (int64_t)(p->nChains - p->nGraftChains) * 2;
		if (w->nAtoms < nGraftAtoms) { startChain = 0; endChain = p->nGraftChains; }
		else { startChain = p->nGraftChains; endChain = p->nChains; }
	} else { startChain = 0; endChain = p->nChains; }                                  /* :245-249 */
	int biased = (p->growthBias[0] != 0 || p->growthBias[1] != 0 || p->growthBias[2] != 0);

	for (int iChain = startChain; iChain < endChain; ++iChain) {                      /* :251 */
		int lastIndex = w->lastIdx[iChain], penIndex = w->penIdx[iChain];             /* :255-256 */
		int iTrial = 0;
		while (iTrial++ < p->maxTrials) {                                             /* :259 */
			double np[3];
			double u = oracle_u_grow(p->seed, (uint32_t)iChain, round, (uint32_t)iTrial);
			w->trials++;
			cone_pos(w->pos + 3 * (size_t)lastIndex, w->pos + 3 * (size_t)penIndex, u, p->trig, w->sinPhi, w->cosPhi, np);   /* :262 */
			if (biased) {   /* extension, see oracle_params.growthBias */
				const double* l = w->pos + 3 * (size_t)lastIndex;
				double sx = np[0] - l[0], sy = np[1] - l[1], sz = np[2] - l[2];
				if ((p->growthBias[0] * sx + p->growthBias[1] * sy) + p->growthBias[2] * sz < 0) continue;
			}
			if (check_collision(w, np, lastIndex)) continue;                          /* :265 */
			add_atom(w, 2, iChain + 1, np, iTrial);                                   /* :267 */
			w->penIdx[iChain] = w->lastIdx[iChain];                                   /* :270-273 */
			w->lastIdx[iChain] = (int32_t)w->nAtoms - 1;
			add_bond(w, w->penIdx[iChain] + 1, w->lastIdx[iChain] + 1);
			w->chainLen[iChain]++;
			break;
		}
		flag = flag && (iTrial >= p->maxTrials);                                      /* :277 */
	}
	return !flag;                                                                     /* :282 */
}

oracle_walk* oracle_walk_run(const oracle_params* p, const double* grafts, int64_t nGrafts, const double* obstacles, int64_t nObst)
{
	oracle_walk* w = (oracle_walk*)calloc(1, sizeof *w);
	w->p = *p;
	w->plook = oracle_lookup_create(p->box[0], p->box[1], p->box[2], p->minDist);     /* sarw.h:88-89 */
	w->lastIdx = (int32_t*)calloc((size_t)p->nChains, sizeof(int32_t));               /* sarw.h:91-93 */
	w->penIdx = (int32_t*)calloc((size_t)p->nChains, sizeof(int32_t));
	w->chainLen = (int32_t*)calloc((size_t)p->nChains, sizeof(int32_t));
	w->grafts = grafts; w->nGrafts = nGrafts;
	for (int64_t i = 0; i < nObst; i++) lookup_add_point(w->plook, obstacles + 3 * i);   /* sarw.cpp:135-142 */
	w->nObstacles = (int)nObst;
	double phi = (180 - 109.5) * M_PI / 180;                                          /* sarw.cpp:165 */
	trig_sincos(p->trig, phi, &w->sinPhi, &w->cosPhi);
	int64_t target = oracle_target_count(p->box, p->targetMassDensity);

	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	w->lastProgressed = 1;
	w->initiated = initiate_chain(w);                                                 /* sarw.cpp:66 */
	if (w->initiated) {
		while (w->nAtoms < target) {                                                  /* sarw.cpp:67 */
			if (p->maxRounds >= 0 && w->rounds >= p->maxRounds) break;
			int ok = propagate_chain(w, (uint32_t)w->rounds);                         /* sarw.cpp:69 */
			if (w->rounds == w->capRounds) {
				w->capRounds = w->capRounds ? 2 * w->capRounds : 2048;
				w->countAfterRound = (int64_t*)realloc(w->countAfterRound, (size_t)w->capRounds * sizeof(int64_t));
			}
			w->countAfterRound[w->rounds++] = w->nAtoms;
			w->lastProgressed = ok;
			if (!ok) break;
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	w->seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	if (w->nAtoms > 0)   /* sarw.cpp:79,359-368: unconditional in main(); unseeded chains have lastIndices = 0 (sarw.h:91) */
		for (int i = 0; i < p->nChains; ++i) w->type[w->lastIdx[i]] = 1;
	return w;
}

void oracle_walk_counts(const oracle_walk* w, int64_t* nAtoms, int64_t* nBonds, int64_t* target, int32_t* initiated, int32_t* rounds,
	int32_t* lastProgressed, double* seconds, uint64_t* trials, uint64_t* seedTrials)
{
	*nAtoms = w->nAtoms; *nBonds = w->nBonds; *target = oracle_target_count(w->p.box, w->p.targetMassDensity);
	*initiated = w->initiated; *rounds = w->rounds; *lastProgressed = w->lastProgressed; *seconds = w->seconds;
	*trials = w->trials; *seedTrials = w->seedTrials;
}

void oracle_walk_get(const oracle_walk* w, double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds, int32_t* chainLengths,
	int64_t* countAfterRound, int32_t* acceptedTrial)
{
	if (xyz) memcpy(xyz, w->pos, (size_t)w->nAtoms * 3 * sizeof(double));
	if (chainID) memcpy(chainID, w->chainID, (size_t)w->nAtoms * sizeof(int32_t));
	if (type) memcpy(type, w->type, (size_t)w->nAtoms * sizeof(int32_t));
	if (acceptedTrial) memcpy(acceptedTrial, w->accTrial, (size_t)w->nAtoms * sizeof(int32_t));
	if (bonds) memcpy(bonds, w->bonds, (size_t)w->nBonds * 2 * sizeof(int64_t));
	if (chainLengths) memcpy(chainLengths, w->chainLen, (size_t)w->p.nChains * sizeof(int32_t));
	if (countAfterRound) memcpy(countAfterRound, w->countAfterRound, (size_t)w->rounds * sizeof(int64_t));
}

void oracle_walk_free(oracle_walk* w)
{
	if (!w) return;
	oracle_lookup_destroy(w->plook);
	free(w->pos); free(w->chainID); free(w->type); free(w->accTrial); free(w->bonds);
	free(w->lastIdx); free(w->penIdx); free(w->chainLen); free(w->countAfterRound);
	free(w);
}

/* ------------------------------------------------------------------ analysis: scripts/rdf.py:13-20, 39-57 */
void oracle_rdf_edges(double rMin, double dr, int32_t nBins, double* edges)
{
	/* numpy.arange for doubles: the first two values are computed, the rest filled as start + k*delta with
	 * delta = second - first (rdf.py:19) */
	volatile double second = rMin + dr;
	double delta = second - rMin;
	for (int32_t k = 0; k <= nBins; k++) edges[k] = rMin + (double)k * delta;
}

void oracle_rdf(const double box[3], const double* xyz, const int32_t* chain, int64_t n, double rMin, double dr, int32_t nBins,
	uint64_t* intra, uint64_t* inter)
{
	double* edges = (double*)malloc(((size_t)nBins + 1) * sizeof(double));
	double Linv[3] = { 1. / box[0], 1. / box[1], 1. / box[2] };          /* rdf.py:14 (1./L) */
	oracle_rdf_edges(rMin, dr, nBins, edges);
	for (int32_t k = 0; k < nBins; k++) { intra[k] = 0; inter[k] = 0; }
	for (int64_t i = 0; i < n; i++) {
		for (int64_t j = 0; j < n; j++) {
			/* rdf.py:49-52: same chain -> rdfPair(chain, chain) (every ordered pair, i == j too); different chains -> only
			 * iChain1 > iChain2 is evaluated and counted twice: every ordered pair contributes one count of (higher - lower) */
			int same = chain[i] == chain[j];
			const double* a1 = (same || chain[i] > chain[j]) ? xyz + 3 * i : xyz + 3 * j;
			const double* a2 = (same || chain[i] > chain[j]) ? xyz + 3 * j : xyz + 3 * i;
			double sum = 0;
			for (int k = 0; k < 3; k++) {
				double d = (a1[k] - a2[k]) * Linv[k];                    /* rdf.py:14 */
				d -= floor(0.5 + d);                                     /* rdf.py:15 */
				d *= box[k];                                             /* rdf.py:16 */
				sum += d * d;                                            /* rdf.py:17 */
			}
			double r = sqrt(sum);
			/* numpy.histogram with explicit edges: edges[k] <= r < edges[k+1], last bin closed on the right */
			if (!(r >= edges[0]) || r > edges[nBins]) continue;
			int32_t lo = 0, hi = nBins;                                  /* largest k with edges[k] <= r */
			while (hi - lo > 1) { int32_t mid = (lo + hi) / 2; if (edges[mid] <= r) lo = mid; else hi = mid; }
			if (r == edges[nBins]) lo = nBins - 1;
			if (same) intra[lo]++; else inter[lo]++;
		}
	}
	free(edges);
}

void oracle_chain_shapes(const double* xyz, const int32_t* chain, int64_t n, int32_t nChains, double* ree2, double* rg2)
{
	double* sum = (double*)calloc((size_t)nChains * 3, sizeof(double));
	int64_t* cnt = (int64_t*)calloc((size_t)nChains, sizeof(int64_t));
	int64_t* first = (int64_t*)malloc((size_t)nChains * sizeof(int64_t));
	int64_t* last = (int64_t*)malloc((size_t)nChains * sizeof(int64_t));
	for (int64_t i = 0; i < n; i++) {
		int32_t c = chain[i] - 1;
		if (!cnt[c]) first[c] = i;
		last[c] = i; cnt[c]++;
		for (int k = 0; k < 3; k++) sum[3 * c + k] += xyz[3 * i + k];
	}
	for (int32_t c = 0; c < nChains; c++) {
		rg2[c] = 0; ree2[c] = 0;
		if (!cnt[c]) continue;
		for (int k = 0; k < 3; k++) sum[3 * c + k] = sum[3 * c + k] / (double)cnt[c];
		double d[3];
		for (int k = 0; k < 3; k++) d[k] = xyz[3 * last[c] + k] - xyz[3 * first[c] + k];
		ree2[c] = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
	}
	for (int64_t i = 0; i < n; i++) {
		int32_t c = chain[i] - 1;
		double d[3];
		for (int k = 0; k < 3; k++) d[k] = xyz[3 * i + k] - sum[3 * c + k];
		rg2[c] += (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
	}
	for (int32_t c = 0; c < nChains; c++) if (cnt[c]) rg2[c] = rg2[c] / (double)cnt[c];
	free(sum); free(cnt); free(first); free(last);
}
