/* sarw_abi.h — C ABI of libsarw_b200.so: the chain-growth hot path of polymer-sarw on one B200 (sm_100a).
 *
 * This is the drop-in boundary. The reference has no plugin/FFI layer; the seam is the two C++
 * classes its main() drives, so each entry point names the reference interface it replaces
 * (file:line under /root/reference):
 *
 *   PeriodicLookup{ctor, addPoint, find}                PeriodicLookup.h:26,37,51
 *   SARW{ctor, initiateChain, propagateChain, terminateChain, addAtom, public state vectors}
 *                                                       sarw.h:77-94,116-126; sarw.cpp:220-368
 *   main()'s growth loop                                sarw.cpp:66-79
 *
 * Conventions: every function returns 0 on success, non-zero on error (sarw_last_error gives the text).
 * "Jammed" (no chain progressed) is NOT an error; it is reported in sarw_stats like the reference's early
 * loop exit (sarw.cpp:69). The handle owns all device memory; the caller owns every host buffer. There is no
 * CPU fallback: without a CUDA device every call fails. One host thread per context.
 */
#ifndef SARW_ABI_H
#define SARW_ABI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sarw_ctx sarw_ctx;

enum { SARW_GRAFT_FILE = 0, SARW_GRAFT_Z = 1, SARW_GRAFT_ZCENTER = 2, SARW_GRAFT_OTHER = 3 };   /* sarw.cpp:342-347 */
enum { SARW_ORDER_ROUNDROBIN = 0, SARW_ORDER_GRAFTED = 1 };                                   /* sarw.cpp:230-249 */
enum { SARW_BOUNDARY_PPP = 0, SARW_BOUNDARY_PPF = 1, SARW_BOUNDARY_PFP = 2, SARW_BOUNDARY_FPP = 3 }; /* sarw.cpp:201-203 */

/* The fields main() reads from the input file (sarw.cpp:22-38) plus the extensions of this build. */
typedef struct sarw_params {
	int64_t nChains;            /* sarw.cpp:22 */
	double boxSize[3];          /* sarw.cpp:23, Angstrom */
	double minDist;             /* sarw.cpp:24 */
	double targetMassDensity;   /* sarw.cpp:30, g/cc */
	int64_t nGraftChains;       /* sarw.cpp:31 */
	int32_t maxTrials;          /* sarw.cpp:35 */
	int32_t graftLoc;           /* sarw.cpp:32, SARW_GRAFT_* */
	int32_t growthOrder;        /* sarw.cpp:33, SARW_ORDER_* */
	int32_t boundary;           /* sarw.cpp:34, SARW_BOUNDARY_* */
	double growthBias[3];       /* README.md:33 only (no reference code): a trial whose step has bias.step < 0 fails */
	uint64_t seed;              /* extension: Philox key; the reference is unseeded (sarw.cpp:16) */
	int32_t device;             /* CUDA device ordinal; < 0 = current device */
	int32_t cellCapacity;       /* 0 = default (3 beads inline per 32-byte cell); 1..2 = test hook forcing overflow handling */
	double cellEdge;            /* 0 = default; minimum cell edge in Angstrom (>= 2.002*minDist is always enforced) */
	int64_t reserveRounds;      /* 0 = auto; growth rounds to pre-allocate bead storage for */
	int32_t launchMode;         /* 0 = auto (one thread-block cluster when nChains <= 256, else cooperative grid); 1 = force grid */
	int32_t maxBlocks;          /* 0 = auto; cap on the thread blocks of the cooperative growth launch, so that several contexts can grow on one
	                               device at the same time (tests run the ranks of a multi-rank melt side by side on one GPU this way) */
	/* Slab decomposition of ONE melt over the GPUs of a node (see "multi-rank" below): with slabRanks >= 2 this context is rank slabRank
	 * of slabRanks and allocates only the cell layers (along z) of its own slab plus slabHalo layers on either side - the memory of the
	 * cell list is partitioned, which is what lets a box larger than one GPU's memory run. 0 / 1 = the whole box (every other use). */
	int32_t slabRank;
	int32_t slabRanks;
	int32_t slabHalo;           /* cell layers kept beyond the slab on either side; 0 = default (8, at least what a round can reach) */
	int32_t reserved0;
} sarw_params;

typedef struct sarw_stats {
	int64_t nBeads;             /* SARW::actualCount(), sarw.h:102 */
	int64_t nBonds;             /* SARW::nBonds(), sarw.h:98 */
	int64_t target;             /* SARW::targetCount(), sarw.h:104 */
	int64_t rounds;             /* propagateChain() calls made so far */
	int32_t initiated;          /* return value of initiateChain(), sarw.cpp:290 */
	int32_t lastProgressed;     /* return value of the last propagateChain(), sarw.cpp:282 */
	uint64_t trials;            /* growth trials evaluated in reference terms (what sarw.cpp:259 would have looped) */
	uint64_t seedTrials;        /* seeding trials in reference terms (sarw.cpp:333) */
	uint64_t queries;           /* overlap queries the device actually evaluated (candidates tested against the cell list) */
	uint64_t conflicts;         /* same-round conflicts re-resolved in chain order */
	uint64_t overflowBeads;     /* beads held in the overflow list because their cell was full */
	double seedMs;              /* device time of seeding (CUDA events on the context's stream) */
	double growMs;              /* device time of all growth launches so far */
	int64_t growLaunches;       /* kernel launches made by growth so far */
	int64_t seedLaunches;
	uint64_t cycPropose;        /* SM cycles block 0 spent per phase of the growth kernel, closing barrier included: */
	uint64_t cycVerify;         /*   propose+insert / verify / conflict clean-up (diagnostics for DESIGN.md) */
	uint64_t cycCleanup;
	int64_t cleanupRounds;      /* rounds that needed the serial conflict clean-up */
	int64_t deadChains;         /* chains proven permanently blocked (skipped from then on; the reference retries them) */
	uint64_t stageFallbacks;    /* hard chain-rounds whose neighbourhood did not fit the shared-memory staging buffer */
	uint64_t arcFallbacks;      /* hard chain-rounds with more blocked arcs than the arc buffer holds (no pre-classification) */
	uint64_t cycSub[8];         /* finer split of cycPropose for warp 0 of block 0: frame, stage, arcs, classify, evaluate, publish, barrier, - */
} sarw_stats;

/* SARW::SARW + PeriodicLookup::PeriodicLookup (sarw.h:88-94, PeriodicLookup.h:26-34): allocates the device cell list. */
int sarw_create(const sarw_params* params, sarw_ctx** out);
void sarw_destroy(sarw_ctx* ctx);
/* text of the last error on this context (or of the last failed sarw_create when ctx == NULL) */
const char* sarw_last_error(const sarw_ctx* ctx);
/* run on a caller-provided CUDA stream (cudaStream_t passed as void*); default is a stream owned by the context */
int sarw_set_stream(sarw_ctx* ctx, void* cudaStream);

/* SARW::addObstacle's effect (sarw.cpp:135-142): points enter the lookup before any bead; host xyz[3*n]. */
int sarw_add_obstacles(sarw_ctx* ctx, const double* xyz, int64_t n);
/* SARW::readGrafts' effect (sarw.cpp:110-119): graft seed positions for chains [0, nGraftChains); host xyz[3*n]. */
int sarw_set_grafts(sarw_ctx* ctx, const double* xyz, int64_t n);

/* SARW::initiateChain (sarw.cpp:290-319): seed two beads per chain, chain i seeing chains < i. */
int sarw_initiate(sarw_ctx* ctx, int32_t* ok);
/* SARW::propagateChain (sarw.cpp:220-283): ONE round. */
int sarw_propagate_round(sarw_ctx* ctx, int32_t* progressed);
/* main()'s loop (sarw.cpp:67-78) on the device: rounds until actualCount >= target, a round makes no progress, or
 * maxRounds more rounds were run (maxRounds < 0: unbounded). target <= 0 means SARW::targetCount(). */
int sarw_grow(sarw_ctx* ctx, int64_t target, int64_t maxRounds, sarw_stats* stats);
/* The same without waiting: sarw_grow_async enqueues seeding (if sarw_initiate was not called yet; main()'s sequence sarw.cpp:66-78)
 * and the growth loop on the context's stream and returns at once; sarw_wait blocks until it finished (ONE synchronisation per melt),
 * continues the loop if the walk needs more rounds than bead storage was reserved for, and fills the stats. One host thread can keep
 * several contexts growing on their own streams this way. No other call on the context in between. */
int sarw_grow_async(sarw_ctx* ctx, int64_t target, int64_t maxRounds);
int sarw_wait(sarw_ctx* ctx, sarw_stats* stats);
/* The large tables of destroyed contexts are kept per device for the next sarw_create of the same size (up to 24 GB); this hands
 * them back to the driver. device < 0: every device. */
int sarw_release_cached_memory(int32_t device);
/* How many contexts like this one can run sarw_grow at the same time on the device, each on its own stream (sarw_set_stream) from
 * its own host thread: a melt of <= 256 chains occupies one thread-block cluster, i.e. a fraction of the SMs (7 melts of
 * BASELINE configs[1] fit a B200). 1 for larger melts, whose cooperative kernel fills the device. No reference counterpart
 * (the reference is serial; one runs several processes, one per host core). */
int32_t sarw_concurrent_capacity(sarw_ctx* ctx);
/* ---- multi-rank: ONE melt over several GPUs, one context per GPU (rank), every context created with identical
 * parameters, obstacles and grafts. State is replicated; only the proposals of a round are partitioned (chain i is
 * proposed by rank i % nranks). Per round: sarw_mr_propose on every rank writes `recordsPerRank` 32-byte records into a
 * caller-owned DEVICE buffer; the caller all-gathers the buffers of all ranks, in rank order, into one device buffer
 * (NCCL over NVLink: torch.distributed.all_gather_into_tensor, ncclAllGather, ...); sarw_mr_finish on every rank applies
 * the other ranks' tentative beads, then verifies, resolves same-round conflicts in chain order and finishes the round
 * exactly like the single-GPU path. Results are identical to sarw_grow for any nranks. After the last round call
 * sarw_mr_commit. The loop is:  while (more) { propose; all-gather; finish(&more); }  commit;                        */
int sarw_mr_setup(sarw_ctx* ctx, int32_t rank, int32_t nranks, int64_t* recordsPerRank, int64_t* recordBytes);
int sarw_mr_more(sarw_ctx* ctx, int32_t* more);       /* same predicate as main()'s loop, sarw.cpp:67-69 */
int sarw_mr_propose(sarw_ctx* ctx, void* d_localRecords);
int sarw_mr_finish(sarw_ctx* ctx, const void* d_allRecords, int32_t* more);
int sarw_mr_commit(sarw_ctx* ctx);
/* Peer-memory variant (one process per GPU on one NVLink/NVSwitch node): the whole loop runs in ONE persistent launch per GPU;
 * each rank stores its records straight into every peer's buffer and the ranks meet on flags in peer memory, so there is no
 * host round trip and no NCCL call per round. After sarw_mr_setup: sarw_mr_peer_export fills a 64-byte CUDA IPC handle; the
 * caller exchanges the handles of all ranks (rank order, 64 bytes each); sarw_mr_peer_connect opens them; then
 * sarw_mr_grow_peer replaces the propose/all-gather/finish loop (call it on every rank at the same time). */
int sarw_mr_peer_export(sarw_ctx* ctx, void* handle64);
int sarw_mr_peer_connect(sarw_ctx* ctx, const void* allHandles);
int sarw_mr_grow_peer(sarw_ctx* ctx, int64_t maxRounds, sarw_stats* stats);
/* Same exchange for contexts that live in ONE process (several ranks side by side on one GPU with sarw_params.maxBlocks, or several
 * GPUs with peer access enabled): after sarw_mr_peer_export, sarw_mr_peer_local_base returns the rank's exchange buffer and
 * sarw_mr_peer_connect_local takes the buffers of all ranks (rank order) instead of IPC handles. */
int sarw_mr_peer_local_base(sarw_ctx* ctx, void** base);
int sarw_mr_peer_connect_local(sarw_ctx* ctx, void* const* bases);
/* Slab mode (sarw_params.slabRanks >= 2): the same calls (sarw_mr_setup with the context's slabRank / slabRanks, peer export / connect,
 * sarw_mr_grow_peer), but the box is decomposed in space. Rank r holds the cell layers of z-slab r plus a halo. Each round a chain is
 * proposed, verified and - on a same-round conflict - re-resolved by the rank whose slab holds the chain's tip; the 32-byte records of
 * the accepted beads are stored into every rank's buffer over NVLink and each rank inserts those that fall into its slab or halo;
 * conflicts are settled by a fixed-point iteration whose results are exchanged the same way. Chain states and bead coordinates stay
 * replicated (72 bytes per bead and rank against ~100 bytes of cell list per bead, which is partitioned), so every rank can download the
 * complete walk; it is bit-identical to the single-GPU walk. sarw_lookup_find_batch is not available on a slab context. */

/* The same for contexts of ONE process (what `sarw --gpus N` uses): ctxs[r] = rank r of n, created with identical parameters (slab
 * contexts: slabRank = r, slabRanks = n) on any devices - several may share a device - and initiated. Enables peer access between the
 * devices, shares the exchange buffers as device pointers and runs sarw_mr_grow_peer on every rank at once, one host thread each.
 * stats (may be NULL) receives rank 0's. */
int sarw_mr_grow_local(sarw_ctx* const* ctxs, int32_t n, int64_t maxRounds, sarw_stats* stats);
/* number of CUDA devices visible to the process (0 if none) */
int32_t sarw_device_count(void);

/* SARW::terminateChain (sarw.cpp:359-368) */
int sarw_terminate(sarw_ctx* ctx);

int sarw_get_stats(sarw_ctx* ctx, sarw_stats* stats);
/* SARW::targetCount (sarw.h:103-104) with a 64-bit volume */
int64_t sarw_target_count(const double boxSize[3], double targetMassDensity);

/* The public state vectors main()'s exporters read (sarw.h:77-83), in the reference's acceptance order:
 * xyz[3*nBeads] (polymerChains[i].pos), chainID[nBeads], type[nBeads], bonds[2*nBonds] (1-based species, listBonds).
 * Any pointer may be NULL. */
int sarw_download(sarw_ctx* ctx, double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds);
/* ---- checks that do not need the whole walk on the host (1e9 beads are 48 GB of state vectors) ----
 * Order-independent checksum of the walk, computed on the device over bead slots (slot = 2*chain + {0,1} for the two seed beads,
 * 2*nChains + round*nChains + chain for the bead a chain accepted in a round): out[0] = sum and out[1] = xor of a 64-bit hash of
 * (slot, bit patterns of x, y, z), out[2] = beads, out[3] = sum of a hash of (bead, slot of its bonded predecessor). A function of the
 * walk alone: one GPU, N GPUs and the CPU oracle give the same four words iff every coordinate and bond is identical
 * (tests/stats.py walk_checksum is the numpy restatement). */
int sarw_walk_checksum(sarw_ctx* ctx, uint64_t out[4]);
/* Beads whose WRAPPED position lies in the axis-aligned box [lo, hi) (0 <= lo < hi <= boxSize), for brute-force verification of samples
 * of a walk too large to download: wrapped xyz[3*cap], ids[cap] (bead slot), prevIds[cap] (slot of the bonded predecessor, -1 for a
 * chain's first bead), chainID[cap] (1-based); arbitrary order. *count = beads found (if > cap the output is truncated). */
int sarw_download_box(sarw_ctx* ctx, const double lo[3], const double hi[3], int64_t cap, int64_t* count, double* xyz, int64_t* ids, int64_t* prevIds, int32_t* chainID);
/* After growth: free the cell table (the analysis and download calls do not need it; at 1e9 beads it is 100 GB). No growth or
 * overlap query on the context afterwards. */
int sarw_release_lookup(sarw_ctx* ctx);
/* chainLengths (sarw.h:83): lengths[nChains] */
int sarw_chain_lengths(sarw_ctx* ctx, int32_t* lengths);
/* actualCount() after each completed round, for the PROGRESS log of sarw.cpp:70-76: counts[min(n, rounds)] */
int sarw_round_counts(sarw_ctx* ctx, int64_t* counts, int64_t n);

/* ---- analysis of the walk on the device (the reference does this off line with scripts/rdf.py) ----
 * scripts/rdf.py:13-20 (rdfPair) + 39-57 (rdfChains), as the integer pair counts they accumulate before normalisation: histograms
 * of the minimum-image distance of all ordered bead pairs, bin k = [rMin + k*dr, rMin + (k+1)*dr) with the last bin closed
 * (numpy.histogram over numpy.arange edges); intra[nBins] = pairs within a chain (both orders and self pairs, rdf.py:49-50),
 * inter[nBins] = pairs of different chains, counted as rdf.py:51-52 does (twice the pairs taken from the higher chain id).
 * Distances beyond the last edge are skipped with a spatial sort + tile bounding boxes, so a cut-off RDF of 1e7 beads is cheap. */
int sarw_rdf(sarw_ctx* ctx, double rMin, double dr, int32_t nBins, uint64_t* intra, uint64_t* inter);
/* end-to-end distance^2 and radius of gyration^2 of every chain from the unwrapped coordinates (sarw.cpp:184 never wraps):
 * ree2[nChains], rg2[nChains]; chains without beads give 0 */
int sarw_chain_shapes(sarw_ctx* ctx, double* ree2, double* rg2);

/* PeriodicLookup::addPoint (PeriodicLookup.h:37-47) for n points; they take the next n lookup indices. */
int sarw_lookup_add_points(sarw_ctx* ctx, const double* xyz, int64_t n);
/* PeriodicLookup::find (PeriodicLookup.h:51-75) for n candidates: hit[i] = any stored point other than ignore[i]
 * (ignored only when >= 0; ignore may be NULL) within minDist under the minimum-image convention. Host buffers. */
int sarw_lookup_find_batch(sarw_ctx* ctx, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit);
/* same with DEVICE pointers (inputs already resident in HBM); runs on the context's stream, does not synchronise */
int sarw_lookup_find_batch_device(sarw_ctx* ctx, const double* d_xyz, const int64_t* d_ignore, int64_t n, uint8_t* d_hit);
/* number of points currently in the lookup (obstacles + beads + added points) */
int64_t sarw_lookup_size(sarw_ctx* ctx);
/* The slab decomposition as pure functions (no device needed): out = {first layer of slab `rank`, one past its last layer, first layer of
 * the rank's window, number of layers in the window (first + count may exceed nLayers: the window wraps around the periodic boundary;
 * count == nLayers: the whole table)}; sarw_slab_owner = the rank whose slab holds `layer`. */
int sarw_slab_layers(int32_t nLayers, int32_t ranks, int32_t rank, int32_t halo, int32_t out[4]);
int32_t sarw_slab_owner(int32_t nLayers, int32_t ranks, int32_t layer);
/* bytes of device memory behind this context's cell list (a slab context holds only its layers: about 1/slabRanks of the table plus the halo) */
int64_t sarw_cell_table_bytes(sarw_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SARW_ABI_H */
