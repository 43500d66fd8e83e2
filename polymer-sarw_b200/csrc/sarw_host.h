// sarw_host.h — what the C-ABI host layer (sarw_abi.cu) and the kernels (sarw_kernels.cu) share.
#pragma once
#include "sarw_device.cuh"

namespace sarw {

// device-resident control block of one context
struct Control {
	unsigned barCount, barGen;        // grid barrier of the persistent growth kernel
	unsigned round;                   // next growth round (= propagateChain() calls made)
	unsigned lastProgressed;          // return value of the last propagateChain(), sarw.cpp:282
	unsigned initiated;               // return value of initiateChain(), sarw.cpp:290
	unsigned seedFailChain;           // first chain whose seeding failed, ID_NONE if none
	unsigned seedDirtyCount;
	unsigned abort;                   // set when a peer stopped answering in the peer-memory multi-rank loop
	unsigned dirty2Count;             // chains put in conflict by the parallel clean-up (second list)
	unsigned nSeeded;                 // chains seeded by initiateChain (all of them, or the index of the first that failed): written once by the seeding kernels
	unsigned ovfCount;                // entries claimed in the overflow table (Tables::ovfCount points here: one read returns it with the rest)
	unsigned long long stageFallbacks, arcFallbacks;   // chain-rounds whose neighbourhood / arcs did not fit the shared-memory buffers
	unsigned long long nBeads;        // SARW::actualCount()
	unsigned long long trials, queries, conflicts, seedTrials;
	unsigned long long cycPropose, cycVerify, cycCleanup;   // SM cycles of block 0 per phase (incl. the closing barrier)
	unsigned cleanupRounds, deadChains;
	unsigned long long cycSub[8];       // finer split of the propose phase of block 0 / warp 0 (diagnostics)
	unsigned int needOrdered[2];        // by round parity: the parallel clean-up left work for the ordered pass (else that pass and its barrier are skipped)
	unsigned long long pathCyc[4], pathCnt[4];   // cluster kernel: propose cycles / chain-rounds by path: dead, fast, fast->hard, hard (diagnostics)
	// slab mode (one melt decomposed over several GPUs)
	unsigned oowCount;                  // probes outside the window of this slab during growth: a protocol error, must stay 0
	unsigned ownCount[2];               // by round parity: chains whose tip lies in this rank's slab
	unsigned cleanResCount, cleanPushCount;   // re-resolutions / pushes of the running clean-up iteration
	unsigned cleanOpen;                 // open entries of the conflict list after the last exchange
	unsigned xseq;                      // clean-up exchanges made so far (tag of the flag handshake)
	unsigned pad1;
	unsigned long long busy[256];       // cluster kernel: per-warp cycles from round start to the first barrier arrive, summed (diagnostics)
};

// slab mode: one message of the clean-up exchange = {nRes, nPush, -, -}, cap 32-byte results, cap chain indices
__host__ __device__ constexpr size_t clean_msg_bytes(uint32_t cap) { return (16 + (size_t)cap * (32 + sizeof(uint32_t)) + 15) & ~(size_t)15; }
// device buffers of a slab context (phase 5 of launch_grow_mr); nullptr for the other multi-rank modes
struct SlabBuffers {
	uint32_t* ownList; uint8_t* ownerOf; void* cleanRes; uint32_t* cleanPush; uint32_t cleanCap; void* peerClean[8];
};

cudaError_t launch_find_batch(const Geom& g, const Tables& T, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit, int numSMs,
	void* scratch, cudaStream_t st);
size_t find_scratch_bytes(const Geom& g, int64_t n);
int cluster_max_active(uint32_t nChains);
cudaError_t prof_read(unsigned long long* out, int n, bool reset);   // -DSARW_PROF builds only
// analysis (sarw_analysis.cu)
size_t rdf_scratch_bytes(int64_t n, size_t* cubBytes);
cudaError_t launch_rdf(const double box[3], const double* d_xyz, const int32_t* d_chain, int64_t n, const double* d_thresh2, int nBins,
	double rMin, double dr, double lastEdge, void* scratch, unsigned long long* d_intra, unsigned long long* d_inter, cudaStream_t st);
cudaError_t launch_chain_shapes(const Tables& T, uint32_t idAtomBase, uint32_t nChains, uint32_t rounds, double* d_ree2, double* d_rg2, cudaStream_t st);
cudaError_t launch_add_points(const Geom& g, const Tables& T, const double* xyz, int64_t n, uint32_t baseId, cudaStream_t st);
cudaError_t launch_seed(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* prev, const double* grafts,
	Control* ctl, uint32_t* dirtyList, int numSMs, cudaStream_t st, int* launches);
cudaError_t grow_max_blocks(int numSMs, int* maxBlocks);
cudaError_t launch_grow(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* origTrial, double* resPos, uint32_t* prev, Control* ctl,
	uint32_t* roundAcc, uint32_t* roundProg, uint32_t* dirtyCount, uint32_t* dirtyList, uint32_t roundEnd, int blocks, int useCluster, cudaStream_t st);
int grow_cluster_max_chains();
cudaError_t launch_grow_mr(const Geom& g, const Tables& T, ChainState* chains, uint32_t* resTrial, uint32_t* origTrial, double* resPos, uint32_t* prev, Control* ctl,
	uint32_t* roundAcc, uint32_t* roundProg, uint32_t* dirtyCount, uint32_t* dirtyList, uint32_t roundEnd, int blocks,
	uint32_t rank, uint32_t ranks, uint32_t phase, void* localRec, const void* allRec, void* const* peerRec, void* const* peerFlags,
	void* const* peerDirty, uint32_t splitVerify, const SlabBuffers* slab, cudaStream_t st);
extern unsigned* g_traceHost[8];   // diagnostics: see launch_grow_mr
cudaError_t download_scan_bytes(uint64_t nSlots, size_t* bytes);
cudaError_t launch_download(const Geom& g, const Tables& T, const uint32_t* prev, const ChainState* chains, uint64_t nSlots, uint32_t nSeeded,
	int terminated, uint32_t* flags, uint32_t* scan, void* cubTemp, size_t cubBytes,
	double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds, cudaStream_t st);
cudaError_t launch_walk_checksum(const Tables& T, const uint32_t* prev, uint32_t idAtomBase, uint64_t nSlots, unsigned long long* out4, int numSMs, cudaStream_t st);
cudaError_t launch_gather_box(const Geom& g, const Tables& T, const uint32_t* prev, uint64_t nSlots, const double lo[3], const double hi[3], uint64_t cap,
	unsigned long long* count, double* xyz, int64_t* ids, int64_t* prevIds, int32_t* chainID, cudaStream_t st);
cudaError_t launch_chain_lengths(const ChainState* chains, uint32_t n, int32_t* out, cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t* p, uint64_t n, uint32_t v, cudaStream_t st);

} // namespace sarw
