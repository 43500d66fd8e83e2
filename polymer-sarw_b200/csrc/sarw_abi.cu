// sarw_abi.cu — host side of the C ABI declared in include/sarw_abi.h: context, device memory, launches.
// No CPU fallback anywhere: every entry point needs a CUDA device and fails with an error text otherwise.
#include "sarw_abi.h"
#include "sarw_host.h"
#include <cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <unordered_map>
#include <thread>

using namespace sarw;

namespace {
thread_local std::string g_createError;

constexpr double AVOGADRO = 6.022e23;   // sarw.h:4
constexpr double BOND_LENGTH = 1.54;    // sarw.h:5
constexpr double ATOM_MASS = 14.0;      // sarw.h:6
}

struct sarw_ctx {
	sarw_params p{};
	Geom g{};
	Tables T{};
	int device = 0, numSMs = 0, growBlocksMax = 0;
	cudaStream_t stream = nullptr; bool ownStream = false;
	alignas(64) CUtensorMap tmapHost;     // tensor map of the cell table (copied into the kernel parameters of every growth launch)
	cudaStream_t allocStream = nullptr;   // all pool allocations / frees go through the device's home stream (cheap reuse)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, evS0 = nullptr, evS1 = nullptr;   // growth launch / seeding, on the context's stream
	std::string err;

	// lookup index space: [0, nObstacles) obstacles / added points, then atom slots
	uint32_t nextId = 0;          // next lookup index for add_points / obstacles
	uint64_t idCapacity = 0;      // lookup indices lpos is allocated for
	bool initiated = false, initiateCalled = false, terminated = false;
	uint32_t nSeeded = 0;
	uint32_t roundCap = 0;        // growth rounds storage is allocated for

	Control* ctl = nullptr;
	Control* hctl = nullptr;      // pinned host mirror of *ctl, refreshed by read_control after every launch sequence (never stale between ABI calls)
	bool lookupReleased = false;  // sarw_release_lookup was called: no cell table any more
	bool seedTimed = true;        // seedMs / initiated / nSeeded have been taken from the events and the mirror
	bool growTimed = true;        // the last growth launch's event pair has been added to growMs
	bool mirrorStale = false;     // work that changes *ctl was enqueued after the last read_control
	bool asyncPending = false;    // sarw_grow_async enqueued work whose result has not been read yet (sarw_wait)
	int64_t asyncTarget = 0, asyncRoundsLeft = 0; uint32_t asyncRoundBefore = 0;
	ChainState* chains = nullptr;
	uint32_t* resTrial = nullptr; uint32_t* origTrial = nullptr; double* resPos = nullptr;
	uint32_t* prev = nullptr; uint64_t prevCap = 0;
	uint32_t *roundAcc = nullptr, *roundProg = nullptr, *dirtyCount = nullptr, *dirtyList = nullptr;
	double* grafts = nullptr; int64_t nGrafts = 0;

	double seedMs = 0, growMs = 0; int64_t growLaunches = 0, seedLaunches = 0;
	uint32_t mrRank = 0, mrRanks = 1; bool mrPending = false;   // multi-rank mode (sarw_mr_*)
	void* xbuf = nullptr; size_t xbufRecBytes = 0, xbufDirtyBytes = 0;               // peer-memory exchange buffer (plain cudaMalloc: IPC-exportable)
	void* peerRec[8] = {}; void* peerFlags[8] = {}; void* peerDirty[8] = {}; bool peersOpen = false, peersLocal = false;
	// slab mode (sarw_params.slabRanks >= 2): the cell table is a virtual address range of the full size of which only this rank's
	// window of z layers is backed by memory (CUDA virtual memory management), so every kernel addresses cells as on a single GPU
	bool slab = false;
	bool cellsVm = false; CUdeviceptr vmBase = 0; size_t vmSize = 0;
	struct VmRange { size_t off, size; CUmemGenericAllocationHandle h; } vmRanges[2] = {}; int nVmRanges = 0;
	size_t windowBytes = 0;
	SlabBuffers sb{};
};

namespace {

int fail(sarw_ctx* c, const char* fmt, ...)
{
	char buf[1024];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	if (c) c->err = buf; else g_createError = buf;
	return 1;
}
#define CU(c, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(c, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static void cache_flush(sarw_ctx* c);

// CUDA virtual memory management through the runtime's driver entry points (no link dependency on libcuda)
struct VmApi {
	CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
	CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
	CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
	CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
	bool ok = false;
};
const VmApi& vm_api()
{
	static VmApi api; static std::once_flag once;
	std::call_once(once, [] {
		auto get = [](const char* name, void** fn) {
			cudaDriverEntryPointQueryResult qr;
			if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); *fn = nullptr; }
			return *fn != nullptr;
		};
		bool ok = get("cuMemGetAllocationGranularity", (void**)&api.getGranularity);
		ok = get("cuMemAddressReserve", (void**)&api.addressReserve) && ok;
		ok = get("cuMemAddressFree", (void**)&api.addressFree) && ok;
		ok = get("cuMemCreate", (void**)&api.create) && ok;
		ok = get("cuMemRelease", (void**)&api.release) && ok;
		ok = get("cuMemMap", (void**)&api.map) && ok;
		ok = get("cuMemUnmap", (void**)&api.unmap) && ok;
		ok = get("cuMemSetAccess", (void**)&api.setAccess) && ok;
		api.ok = ok;
	});
	return api;
}

void vm_free_cells(sarw_ctx* c)
{
	if (!c->cellsVm) return;
	const VmApi& vm = vm_api();
	for (int k = 0; k < c->nVmRanges; k++) { vm.unmap(c->vmBase + c->vmRanges[k].off, c->vmRanges[k].size); vm.release(c->vmRanges[k].h); }
	if (c->vmBase) vm.addressFree(c->vmBase, c->vmSize);
	c->nVmRanges = 0; c->vmBase = 0; c->vmSize = 0; c->cellsVm = false; c->T.cells = nullptr;
}

// The cell table of a slab context: the full table's address range, backed only where this rank's layers [winLo, winLo + winLen) (mod S2)
// lie (a z layer is one contiguous run of S0*S1 cells; the window is one run, or two when it wraps around the periodic boundary).
int vm_alloc_cells_window(sarw_ctx* c, size_t nCells, size_t layerBytes, int winLo, int winLen, int S2)
{
	const VmApi& vm = vm_api();
	if (!vm.ok) return fail(c, "slab mode needs CUDA virtual memory management (cuMemCreate / cuMemMap), which this driver does not offer");
	CUmemAllocationProp prop = {};
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = c->device;
	size_t gran = 0;
	if (vm.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return fail(c, "cuMemGetAllocationGranularity failed");
	const size_t total = ((nCells * sizeof(Cell) + gran - 1) / gran) * gran;
	if (vm.addressReserve(&c->vmBase, total, 0, 0, 0) != CUDA_SUCCESS) return fail(c, "cuMemAddressReserve of %zu bytes failed", total);
	c->vmSize = total; c->cellsVm = true; c->T.cells = reinterpret_cast<Cell*>(c->vmBase);
	// byte ranges of the window, rounded outwards to the mapping granularity and merged when they touch
	size_t lo[2], hi[2]; int n = 0;
	const int first = winLo, last = winLo + winLen;   // last may exceed S2: wraps
	if (last <= S2) { lo[0] = (size_t)first * layerBytes; hi[0] = (size_t)last * layerBytes; n = 1; }
	else { lo[0] = 0; hi[0] = (size_t)(last - S2) * layerBytes; lo[1] = (size_t)first * layerBytes; hi[1] = (size_t)S2 * layerBytes; n = 2; }
	for (int k = 0; k < n; k++) { lo[k] = (lo[k] / gran) * gran; hi[k] = std::min(total, ((hi[k] + gran - 1) / gran) * gran); }
	if (n == 2 && hi[0] >= lo[1]) { hi[0] = hi[1]; n = 1; }
	CUmemAccessDesc acc = {};
	acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	c->windowBytes = 0;
	for (int k = 0; k < n; k++) {
		const size_t size = hi[k] - lo[k];
		CUmemGenericAllocationHandle h;
		CUresult r = vm.create(&h, size, &prop, 0);
		if (r == CUDA_ERROR_OUT_OF_MEMORY) {   // memory parked for reuse or kept by the stream-ordered pool is given back first
			cache_flush(c);
			cudaDeviceSynchronize();
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); else cudaGetLastError();
			r = vm.create(&h, size, &prop, 0);
		}
		if (r != CUDA_SUCCESS) return fail(c, "cuMemCreate of %zu bytes for the cell layers of this slab failed (%d)", size, (int)r);
		if (vm.map(c->vmBase + lo[k], size, 0, h, 0) != CUDA_SUCCESS) { vm.release(h); return fail(c, "cuMemMap failed"); }
		c->vmRanges[c->nVmRanges++] = { lo[k], size, h };
		if (vm.setAccess(c->vmBase + lo[k], size, &acc, 1) != CUDA_SUCCESS) return fail(c, "cuMemSetAccess failed");
		cudaError_t e = cudaMemsetAsync(reinterpret_cast<void*>(c->vmBase + lo[k]), 0xFF, size, c->stream);
		if (e != cudaSuccess) return fail(c, "cudaMemsetAsync of the mapped cell layers: %s", cudaGetErrorString(e));
		c->windowBytes += size;
	}
	return 0;
}

// Device memory. Transient buffers are stream-ordered allocations from the device's default pool. The large long-lived tables
// (cell table, bead storage) of a destroyed context are parked in a per-device cache and handed to the next context that asks for
// the same size: the pool alone made sarw_create erratic (0.4 - 37 ms for the 40 MB cell table of BASELINE configs[1], measured),
// because a warm but fragmented pool remaps physical memory inside cudaMallocAsync.
struct BlockCache {
	std::mutex m;
	std::multimap<size_t, void*> parked;       // size -> block, idle and safe to reuse on any stream
	std::unordered_map<void*, size_t> sizeOf;  // every live large block handed out by dev_alloc
	size_t parkedBytes = 0;
};
static BlockCache g_blocks[64];
constexpr size_t CACHE_MIN_BYTES = (size_t)1 << 20;
constexpr size_t CACHE_MAX_BYTES = (size_t)24 << 30;

static void cache_flush(sarw_ctx* c)
{
	BlockCache& bc = g_blocks[c->device & 63];
	std::lock_guard<std::mutex> lk(bc.m);
	for (auto& kv : bc.parked) cudaFreeAsync(kv.second, c->allocStream);
	bc.parked.clear(); bc.parkedBytes = 0;
	cudaStreamSynchronize(c->allocStream);
}
template <typename T> cudaError_t dev_alloc(sarw_ctx* c, T** p, size_t bytes)
{
	if (bytes >= CACHE_MIN_BYTES) {
		BlockCache& bc = g_blocks[c->device & 63];
		std::lock_guard<std::mutex> lk(bc.m);
		auto it = bc.parked.lower_bound(bytes);
		if (it != bc.parked.end() && it->first <= bytes + bytes / 8) {
			*p = (T*)it->second; bc.sizeOf[it->second] = it->first; bc.parkedBytes -= it->first; bc.parked.erase(it);
			return cudaSuccess;
		}
	}
	cudaError_t e = cudaMallocAsync((void**)p, bytes ? bytes : 16, c->allocStream);
	if (e == cudaErrorMemoryAllocation) {   // give the parked blocks back and try once more
		cudaGetLastError(); cache_flush(c);
		e = cudaMallocAsync((void**)p, bytes ? bytes : 16, c->allocStream);
	}
	if (e == cudaSuccess && c->allocStream != c->stream) e = cudaStreamSynchronize(c->allocStream);   // work runs on a caller's stream
	if (e == cudaSuccess && bytes >= CACHE_MIN_BYTES) {
		BlockCache& bc = g_blocks[c->device & 63];
		std::lock_guard<std::mutex> lk(bc.m);
		bc.sizeOf[(void*)*p] = bytes;
	}
	return e;
}
// stream-ordered free (the caller may still have work queued that uses the block on the context's own home stream)
template <typename T> void dev_free(sarw_ctx* c, T* p)
{
	if (!p) return;
	{	BlockCache& bc = g_blocks[c->device & 63];
		std::lock_guard<std::mutex> lk(bc.m);
		bc.sizeOf.erase((void*)p);
	}
	if (c->allocStream != c->stream && c->stream) cudaStreamSynchronize(c->stream);   // nothing on the work stream may still use it
	cudaFreeAsync((void*)p, c->allocStream);
}
// free of a long-lived table once the context's stream is idle (the caller synchronised it): park it for the next context
template <typename T> void dev_park(sarw_ctx* c, T* p)
{
	if (!p) return;
	{	BlockCache& bc = g_blocks[c->device & 63];
		std::lock_guard<std::mutex> lk(bc.m);
		auto it = bc.sizeOf.find((void*)p);
		if (it != bc.sizeOf.end() && bc.parkedBytes + it->second <= CACHE_MAX_BYTES) {
			bc.parked.emplace(it->second, (void*)p); bc.parkedBytes += it->second; bc.sizeOf.erase(it);
			return;
		}
	}
	dev_free(c, p);
}

uint64_t slots_for_rounds(const sarw_ctx* c, uint64_t rounds) { return 2ull * c->g.nChains + rounds * c->g.nChains; }

// grow (or create) lpos to hold `cap` lookup indices; new entries get the all-ones "never accepted" pattern
int ensure_ids(sarw_ctx* c, uint64_t cap)
{
	if (cap <= c->idCapacity) return 0;
	if (cap >= 0xFFFFFFF0ull) return fail(c, "lookup index space exhausted (%llu indices needed, 32-bit ids)", (unsigned long long)cap);
	uint64_t newCap = std::max<uint64_t>(cap, c->idCapacity + c->idCapacity / 2);
	if (newCap >= 0xFFFFFFF0ull) newCap = 0xFFFFFFEFull;
	double* np = nullptr;
	CU(c, dev_alloc(c, &np, newCap * 3 * sizeof(double)));
	if (c->idCapacity) CU(c, cudaMemcpyAsync(np, c->T.lpos, c->idCapacity * 3 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	CU(c, cudaMemsetAsync(np + 3 * c->idCapacity, 0xFF, (newCap - c->idCapacity) * 3 * sizeof(double), c->stream));
	dev_free(c, c->T.lpos);
	c->T.lpos = np; c->idCapacity = newCap;
	return 0;
}

int ensure_rounds(sarw_ctx* c, uint32_t rounds)
{
	if (rounds <= c->roundCap) return 0;
	uint32_t newCap = std::max<uint32_t>(rounds, c->roundCap + c->roundCap / 2 + 16);
	uint64_t slots = slots_for_rounds(c, newCap);
	if ((uint64_t)c->g.idAtomBase + slots >= 0xFFFFFFF0ull) {
		newCap = rounds; slots = slots_for_rounds(c, newCap);
		if ((uint64_t)c->g.idAtomBase + slots >= 0xFFFFFFF0ull) return fail(c, "bead index space exhausted (32-bit lookup ids)");
	}
	if (ensure_ids(c, c->g.idAtomBase + slots)) return 1;
	// prev[] per slot
	uint32_t* np = nullptr;
	CU(c, dev_alloc(c, &np, slots * sizeof(uint32_t)));
	if (c->prevCap) CU(c, cudaMemcpyAsync(np, c->prev, c->prevCap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
	// per-round counters
	uint32_t* arrs[3]; uint32_t** olds[3] = { &c->roundAcc, &c->roundProg, &c->dirtyCount };
	for (int k = 0; k < 3; k++) {
		CU(c, dev_alloc(c, &arrs[k], (size_t)newCap * sizeof(uint32_t)));
		CU(c, cudaMemsetAsync(arrs[k], 0, (size_t)newCap * sizeof(uint32_t), c->stream));
		if (c->roundCap) CU(c, cudaMemcpyAsync(arrs[k], *olds[k], (size_t)c->roundCap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
	}
	dev_free(c, c->prev);
	for (int k = 0; k < 3; k++) { dev_free(c, *olds[k]); *olds[k] = arrs[k]; }
	c->prev = np; c->prevCap = slots; c->roundCap = newCap;
	return 0;
}

// Pinned control-block mirrors are recycled between contexts (cudaHostAlloc costs far more than a melt's launch).
struct PinnedPool { std::mutex m; std::vector<Control*> idle; };
static PinnedPool g_pinned;
Control* pinned_get()
{
	{	std::lock_guard<std::mutex> lk(g_pinned.m);
		if (!g_pinned.idle.empty()) { Control* p = g_pinned.idle.back(); g_pinned.idle.pop_back(); return p; }
	}
	Control* p = nullptr;
	if (cudaHostAlloc((void**)&p, sizeof(Control), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}
void pinned_put(Control* p)
{
	if (!p) return;
	std::lock_guard<std::mutex> lk(g_pinned.m);
	g_pinned.idle.push_back(p);
}

// ONE synchronisation: the control block (progress, counters, overflow-table fill) lands in the context's pinned mirror. A copy into
// pinned memory is truly asynchronous, so the driver lock is not held while the stream drains (with a pageable destination every other
// host thread's CUDA call queued up behind it for the length of the running kernel: sarw_create stalled 7.5 ms, measured in round 1).
int read_control(sarw_ctx* c, Control* h)
{
	CU(c, cudaMemcpyAsync(c->hctl, c->ctl, sizeof(Control), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	c->mirrorStale = false;
	if (!c->seedTimed) {   // first read after the seeding kernels: what sarw_initiate reports
		float ms = 0; cudaEventElapsedTime(&ms, c->evS0, c->evS1);
		c->seedMs += ms; c->seedTimed = true;
		c->initiated = c->hctl->initiated != 0;
		c->nSeeded = c->hctl->nSeeded;   // not nBeads / 2: with sarw_grow_async the first read happens after the growth loop
	}
	if (!c->growTimed) { float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->growMs += ms; c->growTimed = true; }
	if (h && h != c->hctl) *h = *c->hctl;
	return 0;
}
// the mirror as of now: a read only when something was launched since the last one
int fresh(sarw_ctx* c) { return c->mirrorStale ? read_control(c, nullptr) : 0; }

// after read_control: did the overflow table run out of room?
int check_overflow(sarw_ctx* c)
{
	const uint32_t n = c->hctl->ovfCount;
	if (n > c->T.ovfLimit) return fail(c, "cell overflow table exhausted (%u beads did not fit their cells; raise cellEdge or lower density)", n);
	return 0;
}

void fill_stats(sarw_ctx* c, const Control& h, sarw_stats* s)
{
	memset(s, 0, sizeof *s);
	s->nBeads = (int64_t)h.nBeads;
	s->nBonds = (int64_t)h.nBeads - (int64_t)c->nSeeded;
	if (s->nBonds < 0) s->nBonds = 0;
	s->target = c->g.target;
	s->rounds = h.round;
	s->initiated = (int32_t)h.initiated;
	s->lastProgressed = (int32_t)h.lastProgressed;
	s->trials = h.trials; s->seedTrials = h.seedTrials; s->queries = h.queries; s->conflicts = h.conflicts;
	s->seedMs = c->seedMs; s->growMs = c->growMs; s->growLaunches = c->growLaunches; s->seedLaunches = c->seedLaunches;
	s->cycPropose = h.cycPropose; s->cycVerify = h.cycVerify; s->cycCleanup = h.cycCleanup;
	s->cleanupRounds = h.cleanupRounds; s->deadChains = h.deadChains;
	s->stageFallbacks = h.stageFallbacks; s->arcFallbacks = h.arcFallbacks;
	s->overflowBeads = h.ovfCount;
	for (int k = 0; k < 8; k++) s->cycSub[k] = h.cycSub[k];
}

} // namespace

extern "C" {

// diagnostics (not part of the public header): per-warp busy cycles of the cluster kernel
int sarw_debug_busy(sarw_ctx* c, unsigned long long* out, int n)
{
	if (!c || !out) return 1;
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	for (int k = 0; k < n && k < 256; k++) out[k] = h.busy[k];
	return 0;
}


int64_t sarw_target_count(const double boxSize[3], double targetMassDensity)
{
	// sarw.h:103-104: int vol() truncates; kept, but in 64 bits
	int64_t vol = (int64_t)(boxSize[0] * boxSize[1] * boxSize[2]);
	return (int64_t)nearbyint((vol * targetMassDensity * AVOGADRO * 1e-24) / ATOM_MASS);
}

const char* sarw_last_error(const sarw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int sarw_create(const sarw_params* p, sarw_ctx** out)
{
	if (!p || !out) return fail(nullptr, "sarw_create: null argument");
	*out = nullptr;
	if (p->nChains < 1 || p->nChains > 0x7FFFFFFF) return fail(nullptr, "nChains must be in [1, 2^31)");
	if (!(p->minDist > 0)) return fail(nullptr, "minDist must be positive");
	if (p->maxTrials < 1) return fail(nullptr, "maxTrials must be at least 1");
	for (int k = 0; k < 3; k++)
		if (!(p->boxSize[k] >= 2 * p->minDist))   // the reference needs S = floor(L/minDist) >= 2 (PeriodicLookup.h:16-22,29)
			return fail(nullptr, "boxSize[%d] = %g must be at least 2*minDist = %g", k, p->boxSize[k], 2 * p->minDist);
	if (p->nGraftChains < 0 || p->nGraftChains > p->nChains) return fail(nullptr, "nGraftChains must be in [0, nChains]");
	if (p->slabRanks > 8 || (p->slabRanks >= 2 && (p->slabRank < 0 || p->slabRank >= p->slabRanks)) || p->slabHalo < 0)
		return fail(nullptr, "slab decomposition: need 2 <= slabRanks <= 8, 0 <= slabRank < slabRanks, slabHalo >= 0");

	int nDev = 0;
	cudaError_t e = cudaGetDeviceCount(&nDev);
	if (e != cudaSuccess || nDev == 0) return fail(nullptr, "no CUDA device available (%s); libsarw_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	sarw_ctx* c = new sarw_ctx();
	c->p = *p;
	// SARW_TRACE=1: host-side timestamps of the phases of sarw_create on stderr (diagnostics)
	static const bool trace = getenv("SARW_TRACE") && atoi(getenv("SARW_TRACE")) != 0;
	const auto tr0 = std::chrono::steady_clock::now();
	auto mark = [&](const char* what) { if (trace) fprintf(stderr, "[sarw_create] %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count()); };
	auto bail = [&](int rc) { g_createError = c->err; sarw_destroy(c); return rc; };
	if (p->device >= 0) { if (cudaSetDevice(p->device) != cudaSuccess) { fail(c, "cudaSetDevice(%d) failed", p->device); return bail(1); } }
	cudaGetDevice(&c->device);
	// device facts are queried once per device and process (cudaGetDeviceProperties alone costs milliseconds)
	struct DevInfo { bool valid; int major, minor, numSMs, coop, growBlocksMax; };
	static DevInfo devInfo[64] = {};
	static std::mutex initMutex;                       // contexts may be created from several host threads at once
	std::unique_lock<std::mutex> initLock(initMutex);
	DevInfo& di = devInfo[c->device & 63];
	if (!di.valid) {
		if (cudaDeviceGetAttribute(&di.major, cudaDevAttrComputeCapabilityMajor, c->device) != cudaSuccess
			|| cudaDeviceGetAttribute(&di.minor, cudaDevAttrComputeCapabilityMinor, c->device) != cudaSuccess
			|| cudaDeviceGetAttribute(&di.numSMs, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess
			|| cudaDeviceGetAttribute(&di.coop, cudaDevAttrCooperativeLaunch, c->device) != cudaSuccess) { fail(c, "cudaDeviceGetAttribute failed"); return bail(1); }
		di.growBlocksMax = 0;
		if (di.major >= 10 && grow_max_blocks(di.numSMs, &di.growBlocksMax) != cudaSuccess) { fail(c, "growth kernels cannot be configured: %s", cudaGetErrorString(cudaGetLastError())); return bail(1); }
		di.valid = true;
	}
	initLock.unlock();
	if (di.major < 10) { fail(c, "device %d is sm_%d%d; libsarw_b200 is built for sm_100a only", c->device, di.major, di.minor); return bail(1); }
	c->numSMs = di.numSMs;
	if (!di.coop) { fail(c, "device lacks cooperative launch"); return bail(1); }
	c->growBlocksMax = di.growBlocksMax;

	Geom& g = c->g;
	g.thresh = p->minDist;
	double edgeMin = std::max(2.002 * p->minDist, p->cellEdge > 0 ? p->cellEdge : 2.0);
	double amax = 0;
	for (int k = 0; k < 3; k++) {
		g.L[k] = p->boxSize[k]; g.Linv[k] = 1. / p->boxSize[k];
		int S = (int)floor(p->boxSize[k] / edgeMin); if (S < 1) S = 1;
		g.S[k] = S; g.a[k] = p->boxSize[k] / S; g.ainv[k] = S / p->boxSize[k]; g.sOverL[k] = (double)S;
		// query half-width in cells, strictly below 0.5: a sphere touches at most 2 cells per dimension. The clamp can only bite when
		// S == 1 (a = L < 2.002 minDist), where every bead lives in the one cell that is visited anyway.
		g.reachCells[k] = std::min((p->minDist / g.a[k]) * (1 + 1e-9) + 1e-7, 0.4999999);
		g.qs[k] = (float)(g.a[k] / 65536.0);
		amax = std::max(amax, g.a[k]);
	}
	g.winLo = 0; g.winLen = g.S[2]; g.slabRank = 0; g.slabRanks = 1; g.winStrict = 0; g.tmapZ0 = 0; g.tmapZn = g.S[2];   // the whole table, unless this is one slab of a decomposed melt (below)
	g.cap = (p->cellCapacity >= 1 && p->cellCapacity <= MAX_CAP) ? p->cellCapacity : MAX_CAP;
	{	// fp32 pre-filter bounds. The undecided band covers the half quantisation step of the stored in-cell offsets (edge / 131072 per
		// coordinate), fp32 storage and arithmetic error, all with a wide margin; whatever falls inside is decided in fp64 from lpos.
		double band = 4e-6 * (p->minDist + amax + 2 * BOND_LENGTH) + 1.5 * sqrt(3.0) * amax / 131072.0;
		double lo = std::max(0.0, p->minDist - band), hi = p->minDist + band;
		g.lo2 = nextafterf((float)(lo * lo * (1 - 1e-6)), 0.f);
		g.hi2 = nextafterf((float)(hi * hi * (1 + 1e-6)), INFINITY);
		g.regionReach = p->minDist * (1 + 1e-9) + 1e-6 * amax + band;
		g.arcShrink = 1e-5 + 1.5 * sqrt(3.0) * amax / 131072.0;
	}
	{	double phi = (180 - 109.5) * M_PI / 180, sp, cp;   // sarw.cpp:165
		sincos_portable(phi, sp, cp);
		g.dSinPhi = BOND_LENGTH * sp; g.dCosPhi = BOND_LENGTH * cp;
	}
	for (int k = 0; k < 3; k++) g.bias[k] = p->growthBias[k];
	g.biased = (p->growthBias[0] != 0 || p->growthBias[1] != 0 || p->growthBias[2] != 0) ? 1 : 0;
	g.boundary = p->boundary; g.graftLoc = p->graftLoc; g.growthOrder = p->growthOrder;
	g.maxTrials = p->maxTrials;
	g.key0 = (uint32_t)p->seed; g.key1 = (uint32_t)(p->seed >> 32);
	g.nChains = (uint32_t)p->nChains; g.nGraft = (uint32_t)p->nGraftChains;
	g.idAtomBase = 0; g.seedIgnoreId = ID_NONE;
	g.target = sarw_target_count(p->boxSize, p->targetMassDensity);
	g.nGraftAtoms = (p->nGraftChains * g.target / p->nChains) + (p->nChains - p->nGraftChains) * 2;   // sarw.cpp:232

	{	// one long-lived stream per device, shared by the contexts that do not bring their own (sarw_set_stream): stream-ordered
		// pool allocations are only reused cheaply by the stream that freed them, so per-context streams would pay the
		// driver for fresh memory on every sarw_create
		static cudaStream_t homeStream[64] = {};
		static std::mutex homeMutex;
		std::lock_guard<std::mutex> lk(homeMutex);
		const int d = c->device & 63;
		if (!homeStream[d] && cudaStreamCreateWithFlags(&homeStream[d], cudaStreamNonBlocking) != cudaSuccess) { fail(c, "cudaStreamCreate failed"); return bail(1); }
		c->stream = homeStream[d];
		c->allocStream = homeStream[d];
		c->ownStream = false;
	}
	{	// keep freed blocks in the pool instead of returning them to the driver
		cudaMemPool_t pool; unsigned long long keep = ~0ull;
		if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	}
	mark("device+stream");
	cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1); cudaEventCreate(&c->evS0); cudaEventCreate(&c->evS1);
	mark("events");

	const size_t nCells = (size_t)g.S[0] * g.S[1] * g.S[2];
#define CUB_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(c, "%s: %s", #x, cudaGetErrorString(e_)); return bail(1); } } while (0)
	size_t tableCells = nCells;   // cells backed by memory (sizes the overflow table)
	if (p->slabRanks >= 2) {
		// this rank's slab of z layers plus a halo that covers what a round can reach from a tip in the slab: the staged 4-layer box, the
		// +-1-layer verify probe around the new bead, the candidates of a re-resolution (all within reach + 2 layers of the tip's layer)
		c->slab = true;
		g.slabRank = p->slabRank; g.slabRanks = p->slabRanks;
		const int reach = (int)floor((2.0 + p->minDist + 0.1) / g.a[2]) + 1;
		const int halo = std::max(p->slabHalo > 0 ? p->slabHalo : 8, reach + 2);
		const int lo = slab_lo(g.S[2], g.slabRanks, g.slabRank), hi = slab_lo(g.S[2], g.slabRanks, g.slabRank + 1);
		if (hi - lo < 1) { fail(c, "slab decomposition: %d cell layers cannot be split over %d ranks", g.S[2], g.slabRanks); return bail(1); }
		slab_window(g.S[2], g.slabRanks, g.slabRank, halo, g.winLo, g.winLen);
		if (g.winLen < g.S[2]) {
			if (vm_alloc_cells_window(c, nCells, (size_t)g.S[0] * g.S[1] * sizeof(Cell), g.winLo, g.winLen, g.S[2])) return bail(1);
			tableCells = c->windowBytes / sizeof(Cell);
		}
	}
	if (!c->cellsVm) {
		CUB_(dev_alloc(c, &c->T.cells, nCells * sizeof(Cell)));
		mark("cells alloc");
		CUB_(cudaMemsetAsync(c->T.cells, 0xFF, nCells * sizeof(Cell), c->stream));
	}
	c->T.tmapHost = nullptr; g.useTmap = 0;
	if (g.S[0] >= 4 && g.S[1] >= 4 && g.S[2] >= 4 && !(getenv("SARW_NO_TMAP") && atoi(getenv("SARW_NO_TMAP")))) {
		// 3-D tensor map over the cell table (x in 4-byte elements, 32 per cell; y; z), box = 4 x 4 x 4 cells: a chain's whole
		// neighbourhood is staged by ONE TMA instruction instead of one bulk copy per x-row (each costs ~100 issue cycles).
		// The map travels in the kernel parameters (__grid_constant__): no device copy, nothing to upload in sarw_create.
		typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
			const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
		static EncodeFn encode = nullptr; static bool looked = false; static std::mutex encMutex;
		{	std::lock_guard<std::mutex> lk(encMutex);
			if (!looked) {
				void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
				if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) encode = (EncodeFn)fn;
				else cudaGetLastError();
				looked = true;
			}
		}
		if (encode) {
			// An inner slab's window is one run of layers with unbacked address space before it: its tensor map starts at the window (measured:
			// a tensor copy faults when the map's base address has no memory behind it, even if every element the box touches has).
			// The windows of the first and last slab wrap around the periodic boundary and include the table's first layers.
			if (c->cellsVm && g.winLo + g.winLen <= g.S[2]) { g.tmapZ0 = g.winLo; g.tmapZn = g.winLen; }
			const cuuint64_t dims[3] = { (cuuint64_t)g.S[0] * (sizeof(Cell) / 4), (cuuint64_t)g.S[1], (cuuint64_t)g.tmapZn };
			const cuuint64_t strides[2] = { (cuuint64_t)g.S[0] * sizeof(Cell), (cuuint64_t)g.S[0] * g.S[1] * sizeof(Cell) };
			const cuuint32_t box[3] = { 4 * (cuuint32_t)(sizeof(Cell) / 4), 4, 4 }, es[3] = { 1, 1, 1 };
			if (encode(&c->tmapHost, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, c->T.cells + (size_t)g.tmapZ0 * g.S[0] * g.S[1], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
					CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
				static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is copied into a 128-byte kernel parameter");
				c->T.tmapHost = &c->tmapHost; g.useTmap = 1;
			} else { g.tmapZ0 = 0; g.tmapZn = g.S[2]; }
		}
	}
	{	// overflow hash table: a power of two >= nCells / 16 entries (>= 2^16); at most half of it is ever used
		uint64_t want = std::max<uint64_t>((uint64_t)1 << 16, tableCells / 16);
		if (p->cellCapacity >= 1 && p->cellCapacity < MAX_CAP) want = std::max<uint64_t>(want, tableCells * 4);   // test hook: most beads overflow
		if (c->cellsVm) want += 8 * (uint64_t)g.nChains;   // a slab context keeps the seed beads of the layers it does not hold in this table
		uint64_t size = (uint64_t)1 << 16;
		while (size < want && size < ((uint64_t)1 << 31)) size <<= 1;
		c->T.ovfMask = (uint32_t)(size - 1); c->T.ovfLimit = (uint32_t)(size / 2);
		CUB_(dev_alloc(c, &c->T.ovf, (size_t)size * sizeof(uint4)));
		CUB_(cudaMemsetAsync(c->T.ovf, 0xFF, (size_t)size * sizeof(uint4), c->stream));
	}
	CUB_(dev_alloc(c, &c->ctl, sizeof(Control)));
	c->hctl = pinned_get();
	if (!c->hctl) { fail(c, "cudaHostAlloc of the control-block mirror failed"); return bail(1); }
	memset(c->hctl, 0, sizeof(Control));
	c->hctl->seedFailChain = ID_NONE; c->hctl->lastProgressed = 1;
	CUB_(cudaMemcpyAsync(c->ctl, c->hctl, sizeof(Control), cudaMemcpyHostToDevice, c->stream));
	c->T.ovfCount = &c->ctl->ovfCount;
	c->T.oowCount = &c->ctl->oowCount;
	CUB_(dev_alloc(c, &c->chains, (size_t)g.nChains * sizeof(ChainState)));
	CUB_(cudaMemsetAsync(c->chains, 0, (size_t)g.nChains * sizeof(ChainState), c->stream));
	CUB_(dev_alloc(c, &c->resTrial, (size_t)g.nChains * sizeof(uint32_t)));
	CUB_(cudaMemsetAsync(c->resTrial, 0, (size_t)g.nChains * sizeof(uint32_t), c->stream));
	CUB_(dev_alloc(c, &c->origTrial, (size_t)g.nChains * sizeof(uint32_t)));
	CUB_(cudaMemsetAsync(c->origTrial, 0, (size_t)g.nChains * sizeof(uint32_t), c->stream));
	CUB_(dev_alloc(c, &c->resPos, (size_t)g.nChains * 3 * sizeof(double)));
	CUB_(dev_alloc(c, &c->dirtyList, 3 * (size_t)g.nChains * sizeof(uint32_t)));   // conflict list (n) + second list of the parallel clean-up (<= 2n: independent chains are far apart)
	mark("small allocs");
	CUB_(cudaStreamSynchronize(c->stream));
	mark("sync");
#undef CUB_
	if (c->growBlocksMax < 1) { fail(c, "growth kernel does not fit on the device"); return bail(1); }
	{	// bead storage for the expected number of rounds (grown later if the walk needs more, or if obstacles shift the index space)
		uint64_t wantRounds = p->reserveRounds > 0 ? (uint64_t)p->reserveRounds
			: (uint64_t)std::max<int64_t>(0, (g.target - 2 * (int64_t)g.nChains) / (int64_t)g.nChains) * 9 / 8 + 32;
		if (wantRounds > 0xFFFFFFFFull) wantRounds = 0xFFFFFFFFull;
		if (slots_for_rounds(c, wantRounds) < 0xFFFFFFF0ull && ensure_rounds(c, (uint32_t)std::max<uint64_t>(wantRounds, 1))) return bail(1);
		cudaStreamSynchronize(c->stream);
	}
	mark("bead storage");
	*out = c;
	return 0;
}

void sarw_destroy(sarw_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->stream) {
		if (c->allocStream != c->stream) cudaStreamSynchronize(c->allocStream);   // both streams idle: the blocks may be parked
		if (c->cellsVm) vm_free_cells(c); else dev_park(c, c->T.cells);
		dev_park(c, c->T.lpos); dev_park(c, c->T.ovf);
		dev_free(c, c->sb.ownList); dev_free(c, c->sb.ownerOf); dev_free(c, (char*)c->sb.cleanRes); dev_free(c, c->sb.cleanPush);
		dev_free(c, c->ctl); dev_park(c, c->chains); dev_park(c, c->resTrial); dev_park(c, c->origTrial); dev_park(c, c->resPos); dev_park(c, c->prev);
		dev_park(c, c->roundAcc); dev_park(c, c->roundProg); dev_park(c, c->dirtyCount); dev_park(c, c->dirtyList); dev_park(c, c->grafts);
	}
	if (c->peersOpen && !c->peersLocal) for (uint32_t p = 0; p < c->mrRanks; ++p) if (p != c->mrRank && c->peerRec[p]) cudaIpcCloseMemHandle(c->peerRec[p]);
	if (c->xbuf) cudaFree(c->xbuf);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->evS0) cudaEventDestroy(c->evS0);
	if (c->evS1) cudaEventDestroy(c->evS1);
	if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
	pinned_put(c->hctl);
	delete c;
}

int sarw_set_stream(sarw_ctx* c, void* s)
{
	if (!c) return 1;
	cudaStreamSynchronize(c->stream);
	if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
	c->stream = (cudaStream_t)s; c->ownStream = false;
	return 0;
}

int sarw_lookup_add_points(sarw_ctx* c, const double* xyz, int64_t n)
{
	if (!c) return 1;
	if (n < 0 || (n > 0 && !xyz)) return fail(c, "sarw_lookup_add_points: bad arguments");
	if (c->initiateCalled) return fail(c, "points/obstacles must be added before sarw_initiate");
	if (n == 0) return 0;
	cudaSetDevice(c->device);
	if (ensure_ids(c, (uint64_t)c->nextId + (uint64_t)n)) return 1;
	double* d = nullptr;
	CU(c, dev_alloc(c, &d, (size_t)n * 3 * sizeof(double)));
	CU(c, cudaMemcpyAsync(d, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	cudaError_t e = launch_add_points(c->g, c->T, d, n, c->nextId, c->stream);
	if (e != cudaSuccess) { dev_free(c, d); return fail(c, "add_points failed: %s", cudaGetErrorString(e)); }
	if (read_control(c, nullptr)) { dev_free(c, d); return 1; }
	dev_free(c, d);
	c->nextId += (uint32_t)n;
	return check_overflow(c);
}

int sarw_add_obstacles(sarw_ctx* c, const double* xyz, int64_t n) { return sarw_lookup_add_points(c, xyz, n); }

int64_t sarw_lookup_size(sarw_ctx* c)
{
	if (!c) return -1;
	if (!c->initiateCalled) return c->nextId;
	if (fresh(c)) return -1;
	Control& h = *c->hctl;
	return (int64_t)c->g.idAtomBase + (int64_t)h.nBeads;
}

int sarw_slab_layers(int32_t nLayers, int32_t ranks, int32_t rank, int32_t halo, int32_t out[4])
{
	if (!out || nLayers < 1 || ranks < 1 || rank < 0 || rank >= ranks || halo < 0) return 1;
	int winLo = 0, winLen = nLayers;
	slab_window(nLayers, ranks, rank, halo, winLo, winLen);
	out[0] = slab_lo(nLayers, ranks, rank); out[1] = slab_lo(nLayers, ranks, rank + 1); out[2] = winLo; out[3] = winLen;
	return 0;
}

int32_t sarw_slab_owner(int32_t nLayers, int32_t ranks, int32_t layer)
{
	if (nLayers < 1 || ranks < 1 || layer < 0 || layer >= nLayers) return -1;
	return slab_owner_of(nLayers, ranks, layer);
}

int64_t sarw_cell_table_bytes(sarw_ctx* c)
{
	if (!c || c->lookupReleased) return 0;
	return c->cellsVm ? (int64_t)c->windowBytes : (int64_t)((size_t)c->g.S[0] * c->g.S[1] * c->g.S[2] * sizeof(Cell));
}

int sarw_set_grafts(sarw_ctx* c, const double* xyz, int64_t n)
{
	if (!c) return 1;
	if (n < 0 || (n > 0 && !xyz)) return fail(c, "sarw_set_grafts: bad arguments");
	cudaSetDevice(c->device);
	if (c->grafts) { dev_free(c, c->grafts); c->grafts = nullptr; }
	c->nGrafts = n;
	if (n == 0) return 0;
	CU(c, dev_alloc(c, &c->grafts, (size_t)n * 3 * sizeof(double)));
	CU(c, cudaMemcpyAsync(c->grafts, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return 0;
}

// Batches that revisit cell lines (several candidates per cell line, table larger than L2) are processed in cell-layer order so
// that a line is fetched from HBM once per batch: scratch for the counting sort, or nullptr when ordering would not pay.
static cudaError_t find_batch_ordered(sarw_ctx* c, const double* d_xyz, const int64_t* d_ignore, int64_t n, uint8_t* d_hit)
{
	uint8_t* scratch = nullptr;
	static const bool orderOff = getenv("SARW_FIND_ORDER") && atoi(getenv("SARW_FIND_ORDER")) == 0;   // diagnostics
	const size_t tableBytes = (size_t)c->g.S[0] * c->g.S[1] * c->g.S[2] * sizeof(Cell);
	if (!orderOff && n >= (1 << 18) && c->g.S[2] >= 8 && c->g.S[2] <= 8192 && n < 0xFFFFFFF0ll && tableBytes > ((size_t)96 << 20) && (size_t)n * 8 * sizeof(Cell) > tableBytes) {
		if (dev_alloc(c, &scratch, find_scratch_bytes(c->g, n)) != cudaSuccess) { scratch = nullptr; cudaGetLastError(); }   // no room: unordered
	}
	cudaError_t e = launch_find_batch(c->g, c->T, d_xyz, d_ignore, n, d_hit, c->numSMs, scratch, c->stream);
	if (scratch) dev_free(c, scratch);
	return e;
}

int sarw_lookup_find_batch_device(sarw_ctx* c, const double* d_xyz, const int64_t* d_ignore, int64_t n, uint8_t* d_hit)
{
	if (!c) return 1;
	if (n < 0 || (n > 0 && (!d_xyz || !d_hit))) return fail(c, "sarw_lookup_find_batch_device: bad arguments");
	if (c->cellsVm) return fail(c, "sarw_lookup_find_batch is not available on a slab context (this rank holds only its layers of the cell list)");
	if (c->lookupReleased) return fail(c, "the lookup was released (sarw_release_lookup)");
	cudaSetDevice(c->device);
	CU(c, find_batch_ordered(c, d_xyz, d_ignore, n, d_hit));
	return 0;
}

int sarw_lookup_find_batch(sarw_ctx* c, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit)
{
	if (!c) return 1;
	if (n < 0 || (n > 0 && (!xyz || !hit))) return fail(c, "sarw_lookup_find_batch: bad arguments");
	if (c->cellsVm) return fail(c, "sarw_lookup_find_batch is not available on a slab context (this rank holds only its layers of the cell list)");
	if (c->lookupReleased) return fail(c, "the lookup was released (sarw_release_lookup)");
	if (n == 0) return 0;
	cudaSetDevice(c->device);
	double* dx = nullptr; int64_t* di = nullptr; uint8_t* dh = nullptr;
	int rc = 0;
	auto cleanup = [&]() { dev_free(c, dx); dev_free(c, di); dev_free(c, dh); };
#define CUF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
	CUF(dev_alloc(c, &dx, (size_t)n * 3 * sizeof(double)));
	CUF(dev_alloc(c, &dh, (size_t)n));
	CUF(cudaMemcpyAsync(dx, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if (ignore) {
		CUF(dev_alloc(c, &di, (size_t)n * sizeof(int64_t)));
		CUF(cudaMemcpyAsync(di, ignore, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	}
	CUF(find_batch_ordered(c, dx, di, n, dh));
	CUF(cudaStreamSynchronize(c->stream));
	CUF(cudaMemcpyAsync(hit, dh, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
#undef CUF
	cleanup();
	return 0;
}

// the seeding kernels on the context's stream, no synchronisation
static int initiate_enqueue(sarw_ctx* c)
{
	if (c->initiateCalled) return fail(c, "sarw_initiate was already called on this context");
	cudaSetDevice(c->device);
	Geom& g = c->g;
	if (g.graftLoc == SARW_GRAFT_FILE && g.nGraft > 0 && c->nGrafts < (int64_t)g.nGraft)
		// the reference only logs "Inadequate number of seeds" and then reads out of bounds (sarw.cpp:114-115,343)
		return fail(c, "Inadequate number of graft seeds: %lld given, %u grafted chains", (long long)c->nGrafts, g.nGraft);
	g.idAtomBase = c->nextId;                                  // = nObstacles
	g.seedIgnoreId = c->nextId > 0 ? c->nextId - 1 : ID_NONE;  // sarw.cpp:198,353: find(pos, -1 + nObstacles)
	if (ensure_rounds(c, std::max<uint32_t>(c->roundCap, 1))) return 1;
	if (ensure_ids(c, (uint64_t)g.idAtomBase + slots_for_rounds(c, c->roundCap))) return 1;   // obstacles shift the atom index space

	CU(c, cudaEventRecord(c->evS0, c->stream));
	int launches = 0;
	CU(c, launch_seed(g, c->T, c->chains, c->resTrial, c->prev, c->grafts, c->ctl, c->dirtyList, c->numSMs, c->stream, &launches));
	CU(c, cudaEventRecord(c->evS1, c->stream));
	c->seedLaunches += launches;
	c->initiateCalled = true; c->mirrorStale = true; c->seedTimed = false;
	return 0;
}

int sarw_initiate(sarw_ctx* c, int32_t* ok)
{
	if (!c) return 1;
	if (initiate_enqueue(c)) return 1;
	if (read_control(c, nullptr)) return 1;
	if (ok) *ok = c->initiated ? 1 : 0;
	return check_overflow(c);
}

// one chain per warp inside a single thread-block cluster when the chain count allows it (hardware cluster barrier)
static int use_cluster(const sarw_ctx* c)
{
	if (c->p.launchMode == 1) return 0;
	return (int64_t)c->g.nChains <= (int64_t)grow_cluster_max_chains() ? 1 : 0;
}

static cudaError_t launch_grow_auto(sarw_ctx* c, const Geom& g, uint32_t roundEnd, int blocks)
{
	if (use_cluster(c)) {
		cudaError_t e = launch_grow(g, c->T, c->chains, c->resTrial, c->origTrial, c->resPos, c->prev, c->ctl, c->roundAcc, c->roundProg, c->dirtyCount,
			c->dirtyList, roundEnd, blocks, 1, c->stream);
		if (e == cudaSuccess) return e;
		cudaGetLastError();          // launch-configuration error (nothing ran): fall back to the cooperative grid for good
		c->p.launchMode = 1;
	}
	return launch_grow(g, c->T, c->chains, c->resTrial, c->origTrial, c->resPos, c->prev, c->ctl, c->roundAcc, c->roundProg, c->dirtyCount,
		c->dirtyList, roundEnd, blocks, 0, c->stream);
}

// diagnostics (not in the ABI header): fast-path step cycles of a -DSARW_PROF build
int sarw_debug_prof(unsigned long long* out, int n, int reset) { cudaDeviceSynchronize(); return prof_read(out, n, reset != 0) == cudaSuccess ? 0 : 1; }

// diagnostics (not in the ABI header): out[8] = propose cycles and chain-rounds by path (dead, fast, fast->hard, hard) of the cluster kernel
int sarw_debug_paths(sarw_ctx* c, unsigned long long* out)
{
	if (!c) return 1;
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	for (int k = 0; k < 4; k++) { out[k] = h.pathCyc[k]; out[4 + k] = h.pathCnt[k]; }
	return 0;
}

int sarw_release_cached_memory(int32_t device)
{
	int cur = 0; cudaGetDevice(&cur);
	for (int d = 0; d < 64; d++) {
		if (device >= 0 && d != (device & 63)) continue;
		BlockCache& bc = g_blocks[d];
		std::lock_guard<std::mutex> lk(bc.m);
		if (bc.parked.empty() && device < 0) continue;
		cudaSetDevice(d);
		for (auto& kv : bc.parked) cudaFree(kv.second);   // parked blocks are idle: the synchronous free is safe on any stream
		bc.parked.clear(); bc.parkedBytes = 0;
		// the stream-ordered pool keeps what was freed into it (release threshold: everything); hand that back as well, or a later
		// allocation that does not come from the pool - the cuMemCreate of a slab's cell layers - finds the device full
		// (measured on 8 GPUs: rank 0, which had just grown a 4e8-bead melt on its own, could not back 13 GB of cell layers)
		cudaDeviceSynchronize();
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); else cudaGetLastError();
	}
	cudaSetDevice(cur);
	return 0;
}

int32_t sarw_concurrent_capacity(sarw_ctx* c)
{
	if (!c) return 0;
	cudaSetDevice(c->device);
	// small melts grow in ONE thread-block cluster (a fraction of the SMs): several contexts can grow at the same time, each on its
	// own stream from its own host thread. The cooperative grid kernel of larger melts fills the device by itself.
	if (!use_cluster(c)) return 1;
	const int n = cluster_max_active(c->g.nChains);
	return n < 1 ? 1 : n;
}

// CTAs of a cooperative growth launch: enough for one warp per chain, at most what fits the device (or the caller's cap)
static int grow_blocks(const sarw_ctx* c, int64_t chains)
{
	int64_t cap = c->growBlocksMax;
	if (c->p.maxBlocks > 0 && c->p.maxBlocks < cap) cap = c->p.maxBlocks;
	const int blocks = (int)std::min<int64_t>(cap, (chains + 7) / 8);
	return blocks < 1 ? 1 : blocks;
}

// one growth launch on the context's stream covering as many rounds as bead storage is allocated for; no synchronisation
static int grow_enqueue(sarw_ctx* c, const Geom& g, uint32_t roundNow, int64_t roundsLeft)
{
	if (c->lookupReleased) return fail(c, "the lookup was released (sarw_release_lookup): no more growth on this context");
	if (c->slab) return fail(c, "this context holds one slab of a decomposed box: grow it with sarw_mr_grow_peer / sarw_mr_grow_local together with the other slabs");
	if (roundNow >= c->roundCap && ensure_rounds(c, roundNow + 1)) return 1;
	uint32_t roundEnd = c->roundCap;
	if ((int64_t)(roundEnd - roundNow) > roundsLeft) roundEnd = roundNow + (uint32_t)roundsLeft;
	CU(c, cudaEventRecord(c->ev0, c->stream));
	CU(c, launch_grow_auto(c, g, roundEnd, grow_blocks(c, g.nChains)));
	CU(c, cudaEventRecord(c->ev1, c->stream));
	c->mirrorStale = true; c->growTimed = false; c->growLaunches += 1;
	return 0;
}

// main()'s loop: launch, read the control block (ONE synchronisation per launch), repeat only if round storage ran out
static int grow_loop(sarw_ctx* c, const Geom& g, int64_t roundsLeft)
{
	Control& h = *c->hctl;
	// main() only enters the loop when initiateChain succeeded (sarw.cpp:66)
	while (c->initiated && roundsLeft > 0 && (int64_t)h.nBeads < g.target && h.lastProgressed) {
		const uint32_t before = h.round;
		if (grow_enqueue(c, g, before, roundsLeft)) return 1;
		if (read_control(c, nullptr)) return 1;
		roundsLeft -= (int64_t)(h.round - before);
		if (check_overflow(c)) return 1;
		if (h.round == before) break;
	}
	return 0;
}

static int grow_impl(sarw_ctx* c, int64_t target, int64_t maxRounds, sarw_stats* stats)
{
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called before growing");
	if (c->asyncPending) return fail(c, "sarw_wait must be called before another growth call");
	cudaSetDevice(c->device);
	Geom g = c->g;
	if (target > 0) g.target = target;
	if (fresh(c)) return 1;
	if (grow_loop(c, g, maxRounds < 0 ? INT64_MAX : maxRounds)) return 1;
	if (stats) fill_stats(c, *c->hctl, stats);
	return 0;
}

int sarw_grow(sarw_ctx* c, int64_t target, int64_t maxRounds, sarw_stats* stats)
{
	if (!c) return 1;
	return grow_impl(c, target, maxRounds, stats);
}

int sarw_grow_async(sarw_ctx* c, int64_t target, int64_t maxRounds)
{
	if (!c) return 1;
	if (c->asyncPending) return fail(c, "sarw_wait must be called before another growth call");
	cudaSetDevice(c->device);
	// main()'s sequence (sarw.cpp:66-78): initiateChain, then the loop. Whether seeding succeeded is checked by the growth kernel
	// itself (it reads `initiated` from the control block), so nothing has to come back to the host in between.
	if (!c->initiateCalled) { if (initiate_enqueue(c)) return 1; }
	else if (fresh(c)) return 1;
	Geom g = c->g;
	if (target > 0) g.target = target;
	c->asyncTarget = g.target; c->asyncRoundsLeft = maxRounds < 0 ? INT64_MAX : maxRounds; c->asyncRoundBefore = c->hctl->round;
	if (c->asyncRoundsLeft > 0 && grow_enqueue(c, g, c->asyncRoundBefore, c->asyncRoundsLeft)) return 1;
	c->asyncPending = true;
	return 0;
}

int sarw_wait(sarw_ctx* c, sarw_stats* stats)
{
	if (!c) return 1;
	cudaSetDevice(c->device);
	if (c->asyncPending) {
		c->asyncPending = false;
		if (read_control(c, nullptr)) return 1;
		if (check_overflow(c)) return 1;
		Control& h = *c->hctl;
		Geom g = c->g; g.target = c->asyncTarget;
		const int64_t left = c->asyncRoundsLeft - (int64_t)(h.round - c->asyncRoundBefore);
		// the launch stops at the last round bead storage was allocated for: continue (synchronously) if the walk needs more
		if (h.round > c->asyncRoundBefore && grow_loop(c, g, left)) return 1;
	} else if (fresh(c)) return 1;
	if (stats) fill_stats(c, *c->hctl, stats);
	return 0;
}

int sarw_propagate_round(sarw_ctx* c, int32_t* progressed)
{
	if (!c) return 1;
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called before sarw_propagate_round");
	if (c->asyncPending) return fail(c, "sarw_wait must be called before another growth call");
	// one unconditional round, like a direct call of SARW::propagateChain
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	if (!c->initiated) return fail(c, "initiateChain failed; the reference never propagates in that case (sarw.cpp:66)");
	Control& h = *c->hctl;
	Geom g = c->g;
	g.target = INT64_MAX;
	h.lastProgressed = 1;   // one unconditional round
	CU(c, cudaMemcpyAsync(&c->ctl->lastProgressed, &h.lastProgressed, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));   // source is pinned
	if (grow_enqueue(c, g, h.round, 1)) return 1;
	if (read_control(c, nullptr)) return 1;
	if (progressed) *progressed = (int32_t)h.lastProgressed;
	return check_overflow(c);
}

// ------------------------------------------------------------------ one melt over several GPUs (include/sarw_abi.h, "multi-rank")
int sarw_mr_setup(sarw_ctx* c, int32_t rank, int32_t nranks, int64_t* recordsPerRank, int64_t* recordBytes)
{
	if (!c) return 1;
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, "sarw_mr_setup: rank %d of %d", rank, nranks);
	if (c->slab && (rank != c->p.slabRank || nranks != c->p.slabRanks)) return fail(c, "sarw_mr_setup: this context was created as slab %d of %d", c->p.slabRank, c->p.slabRanks);
	c->mrRank = (uint32_t)rank; c->mrRanks = (uint32_t)nranks;
	if (recordsPerRank) *recordsPerRank = ((int64_t)c->g.nChains + nranks - 1) / nranks;
	if (recordBytes) *recordBytes = 32;
	return 0;
}

static int mr_launch(sarw_ctx* c, uint32_t phase, void* localRec, const void* allRec, bool readBack)
{
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called before the multi-rank rounds");
	if (c->slab) return fail(c, "slab contexts grow with sarw_mr_grow_peer (the exchange runs inside the kernel)");
	cudaSetDevice(c->device);
	if (phase != 2u && fresh(c)) return 1;   // phase 2 follows phase 1 of the same round: the round number in the mirror is still right
	Control& h = *c->hctl;
	if (h.round >= c->roundCap && ensure_rounds(c, h.round + 1)) return 1;
	CU(c, cudaEventRecord(c->ev0, c->stream));
	CU(c, launch_grow_mr(c->g, c->T, c->chains, c->resTrial, c->origTrial, c->resPos, c->prev, c->ctl, c->roundAcc, c->roundProg, c->dirtyCount, c->dirtyList,
		h.round + 1, grow_blocks(c, c->g.nChains), c->mrRank, c->mrRanks, phase, localRec, allRec, nullptr, nullptr, nullptr, 0u, nullptr, c->stream));
	CU(c, cudaEventRecord(c->ev1, c->stream));
	c->growLaunches += 1; c->mirrorStale = true;
	if (readBack) {
		c->growTimed = false;
		if (read_control(c, nullptr)) return 1;
		return check_overflow(c);
	}
	return 0;
}

int sarw_mr_more(sarw_ctx* c, int32_t* more)
{
	if (!c || !more) return 1;
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	*more = (c->initiated && (int64_t)h.nBeads < c->g.target && h.lastProgressed) ? 1 : 0;
	return 0;
}

int sarw_mr_propose(sarw_ctx* c, void* d_localRecords)
{
	if (!c || !d_localRecords) return 1;
	if (mr_launch(c, 1u, d_localRecords, nullptr, false)) return 1;
	c->mrPending = true;
	return 0;
}

int sarw_mr_finish(sarw_ctx* c, const void* d_allRecords, int32_t* more)
{
	if (!c || !d_allRecords) return 1;
	if (mr_launch(c, 2u, nullptr, d_allRecords, true)) return 1;
	Control& h = *c->hctl;
	if (more) *more = (c->initiated && (int64_t)h.nBeads < c->g.target && h.lastProgressed) ? 1 : 0;
	return 0;
}

int sarw_mr_commit(sarw_ctx* c)
{
	if (!c) return 1;
	if (!c->mrPending) return 0;
	if (mr_launch(c, 3u, nullptr, nullptr, true)) return 1;
	c->mrPending = false;
	return 0;
}

// ---- peer-memory variant: the whole loop in one persistent launch per GPU, exchange over NVLink inside the kernel
int sarw_mr_peer_export(sarw_ctx* c, void* handle64)
{
	if (!c || !handle64) return 1;
	if (c->mrRanks < 2 || c->mrRanks > 8) return fail(c, "sarw_mr_peer_export: call sarw_mr_setup with 2..8 ranks first");
	cudaSetDevice(c->device);
	const size_t K = ((size_t)c->g.nChains + c->mrRanks - 1) / c->mrRanks;
	c->xbufRecBytes = c->slab ? 2 * (size_t)c->g.nChains * 32 : 2 * (size_t)c->mrRanks * K * 32;   // slab mode: one record slot per chain
	const size_t dirtyBytes = 2 * (size_t)c->mrRanks * (1 + K) * sizeof(uint32_t);   // conflict-list inboxes
	c->xbufDirtyBytes = (dirtyBytes + 15) & ~(size_t)15;
	size_t cleanBytes = 0;
	if (c->slab) {
		c->sb.cleanCap = (uint32_t)std::min<uint64_t>(65536, std::max<uint64_t>(1024, c->g.nChains / 16));
		cleanBytes = 2 * (size_t)c->mrRanks * clean_msg_bytes(c->sb.cleanCap);
		if (!c->sb.ownList) {
			CU(c, dev_alloc(c, &c->sb.ownList, (size_t)c->g.nChains * sizeof(uint32_t)));
			CU(c, dev_alloc(c, &c->sb.ownerOf, (size_t)c->g.nChains));
			CU(c, dev_alloc(c, (char**)&c->sb.cleanRes, (size_t)c->sb.cleanCap * 32));
			CU(c, dev_alloc(c, &c->sb.cleanPush, (size_t)c->sb.cleanCap * sizeof(uint32_t)));
		}
	}
	if (!c->xbuf) {
		CU(c, cudaMalloc(&c->xbuf, c->xbufRecBytes + 256 + c->xbufDirtyBytes + cleanBytes));
		CU(c, cudaMemset(c->xbuf, 0, c->xbufRecBytes + 256 + c->xbufDirtyBytes + cleanBytes));
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
	CU(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, c->xbuf));
	return 0;
}

int sarw_mr_peer_connect(sarw_ctx* c, const void* allHandles)
{
	if (!c || !allHandles) return 1;
	if (!c->xbuf) return fail(c, "sarw_mr_peer_connect: call sarw_mr_peer_export first");
	cudaSetDevice(c->device);
	for (uint32_t p = 0; p < c->mrRanks; ++p) {
		void* base = c->xbuf;
		if (p != c->mrRank) {
			cudaIpcMemHandle_t h; memcpy(&h, (const char*)allHandles + 64 * p, 64);
			CU(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
		}
		c->peerRec[p] = base; c->peerFlags[p] = (char*)base + c->xbufRecBytes; c->peerDirty[p] = (char*)base + c->xbufRecBytes + 256;
		c->sb.peerClean[p] = (char*)base + c->xbufRecBytes + 256 + c->xbufDirtyBytes;
	}
	c->peersOpen = true;
	return 0;
}

// contexts of ONE process (ranks emulated on one GPU, or several GPUs of one process with peer access enabled): the exchange
// buffers are plain device pointers, no IPC. bases[r] = sarw_mr_peer_local_base of rank r's context.
int sarw_mr_peer_local_base(sarw_ctx* c, void** base)
{
	if (!c || !base) return 1;
	if (!c->xbuf) return fail(c, "sarw_mr_peer_local_base: call sarw_mr_peer_export first");
	*base = c->xbuf;
	return 0;
}

int sarw_mr_peer_connect_local(sarw_ctx* c, void* const* bases)
{
	if (!c || !bases) return 1;
	if (!c->xbuf) return fail(c, "sarw_mr_peer_connect_local: call sarw_mr_peer_export first");
	for (uint32_t p = 0; p < c->mrRanks; ++p) {
		char* base = (char*)(p == c->mrRank ? c->xbuf : bases[p]);
		if (!base) return fail(c, "sarw_mr_peer_connect_local: no buffer for rank %u", p);
		c->peerRec[p] = base; c->peerFlags[p] = base + c->xbufRecBytes; c->peerDirty[p] = base + c->xbufRecBytes + 256;
		c->sb.peerClean[p] = base + c->xbufRecBytes + 256 + c->xbufDirtyBytes;
	}
	c->peersOpen = true; c->peersLocal = true;
	return 0;
}

int sarw_mr_grow_peer(sarw_ctx* c, int64_t maxRounds, sarw_stats* stats)
{
	if (!c) return 1;
	if (!c->peersOpen) return fail(c, "sarw_mr_grow_peer: peers are not connected");
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called before growing");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	int64_t roundsLeft = maxRounds < 0 ? INT64_MAX : maxRounds;
	while (c->initiated && roundsLeft > 0 && (int64_t)h.nBeads < c->g.target && h.lastProgressed) {
		if (h.round >= c->roundCap && ensure_rounds(c, h.round + 1)) return 1;
		uint32_t roundEnd = c->roundCap;
		if ((int64_t)(roundEnd - h.round) > roundsLeft) roundEnd = h.round + (uint32_t)roundsLeft;
		CU(c, cudaEventRecord(c->ev0, c->stream));
		CU(c, launch_grow_mr(c->g, c->T, c->chains, c->resTrial, c->origTrial, c->resPos, c->prev, c->ctl, c->roundAcc, c->roundProg, c->dirtyCount, c->dirtyList,
			roundEnd, grow_blocks(c, (int64_t)c->g.nChains / c->mrRanks), c->mrRank, c->mrRanks, c->slab ? 5u : 4u, nullptr, nullptr, c->peerRec, c->peerFlags, c->peerDirty,
			(c->p.launchMode & 2) ? 0u : 1u, c->slab ? &c->sb : nullptr, c->stream));   // launchMode bit 1: keep the verification replicated (diagnostics)
		CU(c, cudaEventRecord(c->ev1, c->stream));
		c->mirrorStale = true; c->growTimed = false; c->growLaunches += 1;
		const uint32_t before = h.round;
		if (read_control(c, nullptr)) {
			if (unsigned* t = g_traceHost[c->mrRank & 7]) c->err += " [trace: round " + std::to_string(t[0] >> 8) + ", stage " + std::to_string(t[0] & 255u) + "]";
			return 1;
		}
		roundsLeft -= (int64_t)(h.round - before);
		if (h.abort == 2u || h.abort == 3u) return fail(c, "multi-rank growth aborted: more same-round conflicts than the exchange buffers hold (round %u)", h.round);
		if (h.abort) return fail(c, "multi-rank growth aborted: a peer stopped answering (round %u)", h.round);
		if (h.oowCount) return fail(c, "slab decomposition: %u cell probes fell outside this rank's window of layers (halo too small: raise slabHalo)", h.oowCount);
		if (check_overflow(c)) return 1;
		if (h.round == before) break;
	}
	if (stats) fill_stats(c, h, stats);
	return 0;
}

int32_t sarw_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int sarw_mr_grow_local(sarw_ctx* const* ctxs, int32_t n, int64_t maxRounds, sarw_stats* stats)
{
	if (!ctxs || n < 1 || n > 8) return 1;
	for (int32_t r = 0; r < n; r++) if (!ctxs[r]) return 1;
	sarw_ctx* c0 = ctxs[0];
	if (n == 1) return sarw_grow(c0, 0, maxRounds, stats);
	// contexts that share a device must fit it together: the cooperative grids of all ranks have to be resident at the same time
	for (int32_t r = 0; r < n; r++) {
		int sharing = 0;
		for (int32_t q = 0; q < n; q++) sharing += ctxs[q]->device == ctxs[r]->device;
		if (sharing > 1) {
			const int cap = std::max(1, ctxs[r]->growBlocksMax / sharing);
			if (ctxs[r]->p.maxBlocks <= 0 || ctxs[r]->p.maxBlocks > cap) ctxs[r]->p.maxBlocks = cap;
		}
	}
	for (int32_t a = 0; a < n; a++)
		for (int32_t b = 0; b < n; b++) {
			if (ctxs[a]->device == ctxs[b]->device) continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device);
			if (!can) return fail(c0, "sarw_mr_grow_local: device %d cannot access the memory of device %d", ctxs[a]->device, ctxs[b]->device);
			cudaSetDevice(ctxs[a]->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c0, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctxs[a]->device, ctxs[b]->device, cudaGetErrorString(e));
			cudaGetLastError();
		}
	// every rank's persistent kernel must run at the same time: contexts that still share a stream (the device's home stream) get their own
	for (int32_t r = 1; r < n; r++) {
		bool shared = false;
		for (int32_t q = 0; q < r; q++) shared = shared || ctxs[q]->stream == ctxs[r]->stream;
		if (!shared) continue;
		cudaSetDevice(ctxs[r]->device);
		cudaStreamSynchronize(ctxs[r]->stream);
		cudaStream_t s = nullptr;
		if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return fail(c0, "cudaStreamCreate failed");
		ctxs[r]->stream = s; ctxs[r]->ownStream = true;
	}
	void* bases[8] = {};
	for (int32_t r = 0; r < n; r++) {
		unsigned char handle[64];
		if (sarw_mr_setup(ctxs[r], r, n, nullptr, nullptr) || sarw_mr_peer_export(ctxs[r], handle) || sarw_mr_peer_local_base(ctxs[r], &bases[r])) {
			if (r) c0->err = ctxs[r]->err;
			return 1;
		}
	}
	for (int32_t r = 0; r < n; r++)
		if (sarw_mr_peer_connect_local(ctxs[r], bases)) { if (r) c0->err = ctxs[r]->err; return 1; }
	for (int32_t r = 0; r < n; r++) { cudaSetDevice(ctxs[r]->device); cudaStreamSynchronize(ctxs[r]->stream); }
	int rc[8] = {}; sarw_stats st[8];
	std::vector<std::thread> th;
	for (int32_t r = 0; r < n; r++) th.emplace_back([&, r] { rc[r] = sarw_mr_grow_peer(ctxs[r], maxRounds, &st[r]); });
	for (auto& t : th) t.join();
	for (int32_t r = 0; r < n; r++) if (rc[r]) { if (r) c0->err = "rank " + std::to_string(r) + ": " + ctxs[r]->err; return 1; }
	if (stats) *stats = st[0];
	return 0;
}

int sarw_terminate(sarw_ctx* c)
{
	if (!c) return 1;
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called before sarw_terminate");
	c->terminated = true;   // applied when atoms are gathered: type of each chain's last bead := 1 (sarw.cpp:365)
	return 0;
}

int sarw_get_stats(sarw_ctx* c, sarw_stats* s)
{
	if (!c || !s) return 1;
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	fill_stats(c, h, s);
	return 0;
}

int sarw_download(sarw_ctx* c, double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds)
{
	if (!c) return 1;
	if (!c->initiateCalled) return fail(c, "nothing to download before sarw_initiate");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	const uint64_t N = h.nBeads, nBonds = N > c->nSeeded ? N - c->nSeeded : 0;
	if (N == 0) return 0;
	const uint64_t nSlots = slots_for_rounds(c, h.round);
	uint32_t *flags = nullptr, *scan = nullptr; void* tmp = nullptr; size_t tmpBytes = 0;
	double* dx = nullptr; int32_t *dc = nullptr, *dt = nullptr; int64_t* db = nullptr;
	int rc = 0;
	auto cleanup = [&]() { dev_free(c, flags); dev_free(c, scan); dev_free(c, (char*)tmp); dev_free(c, dx); dev_free(c, dc); dev_free(c, dt); dev_free(c, db); };
#define CUF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
	CUF(dev_alloc(c, &flags, nSlots * sizeof(uint32_t)));
	CUF(dev_alloc(c, &scan, nSlots * sizeof(uint32_t)));
	CUF(download_scan_bytes(nSlots, &tmpBytes));
	CUF(dev_alloc(c, (char**)&tmp, tmpBytes ? tmpBytes : 16));
	if (xyz) CUF(dev_alloc(c, &dx, N * 3 * sizeof(double)));
	if (chainID) CUF(dev_alloc(c, &dc, N * sizeof(int32_t)));
	if (type) CUF(dev_alloc(c, &dt, N * sizeof(int32_t)));
	if (bonds && nBonds) CUF(dev_alloc(c, &db, nBonds * 2 * sizeof(int64_t)));
	CUF(launch_download(c->g, c->T, c->prev, c->chains, nSlots, c->nSeeded, c->terminated ? 1 : 0, flags, scan, tmp, tmpBytes, dx, dc, dt, db, c->stream));
	CUF(cudaStreamSynchronize(c->stream));   // see read_control
	if (dx) CUF(cudaMemcpyAsync(xyz, dx, N * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if (dc) CUF(cudaMemcpyAsync(chainID, dc, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
	if (dt) CUF(cudaMemcpyAsync(type, dt, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
	if (db) CUF(cudaMemcpyAsync(bonds, db, nBonds * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
#undef CUF
	cleanup();
	return 0;
}

// ---- analysis of the finished walk (scripts/rdf.py)
// smallest double x >= 0 whose correctly rounded square root is >= e (strict: > e): pairs are binned on squared distances
static double sqrt_threshold(double e, bool strict)
{
	if (e < 0 || (!strict && e == 0)) return 0;
	auto reaches = [&](double x) { const double r = sqrt(x); return strict ? r > e : r >= e; };
	double x = e * e;
	while (x > 0 && reaches(nextafter(x, 0.0))) x = nextafter(x, 0.0);
	while (!reaches(x)) x = nextafter(x, INFINITY);
	return x;
}

int sarw_rdf(sarw_ctx* c, double rMin, double dr, int32_t nBins, uint64_t* intra, uint64_t* inter)
{
	if (!c) return 1;
	if (!intra || !inter || nBins < 1 || nBins > 4096 || !(dr > 0) || !std::isfinite(rMin)) return fail(c, "sarw_rdf: bad arguments (1 <= nBins <= 4096, dr > 0)");
	if (!c->initiateCalled) return fail(c, "nothing to analyse before sarw_initiate");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	const uint64_t N = h.nBeads;
	for (int32_t k = 0; k < nBins; k++) { intra[k] = 0; inter[k] = 0; }
	if (N == 0) return 0;
	if (N > 0x7FFFFFF0ull) return fail(c, "sarw_rdf: more than 2^31 beads are not supported");
	// bin edges as numpy.arange fills them (rdf.py:19): start + k*((start + dr) - start)
	std::vector<double> t2((size_t)nBins + 1);
	volatile double second = rMin + dr;
	const double delta = second - rMin;
	double lastEdge = 0;
	for (int32_t k = 0; k <= nBins; k++) { const double e = rMin + (double)k * delta; t2[(size_t)k] = sqrt_threshold(e, k == nBins); lastEdge = e; }
	if (lastEdge < 0) return 0;
	const uint64_t nSlots = slots_for_rounds(c, h.round);
	uint32_t *flags = nullptr, *scan = nullptr; char *tmp = nullptr, *scratch = nullptr; size_t tmpBytes = 0;
	double *dx = nullptr, *dT = nullptr; int32_t* dc = nullptr; unsigned long long* dh = nullptr;
	int rc = 0;
	auto cleanup = [&]() { dev_free(c, flags); dev_free(c, scan); dev_free(c, tmp); dev_free(c, scratch); dev_free(c, dx); dev_free(c, dT); dev_free(c, dc); dev_free(c, dh); };
#define CUF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
	CUF(dev_alloc(c, &flags, nSlots * sizeof(uint32_t)));
	CUF(dev_alloc(c, &scan, nSlots * sizeof(uint32_t)));
	CUF(download_scan_bytes(nSlots, &tmpBytes));
	CUF(dev_alloc(c, &tmp, tmpBytes ? tmpBytes : 16));
	CUF(dev_alloc(c, &dx, N * 3 * sizeof(double)));
	CUF(dev_alloc(c, &dc, N * sizeof(int32_t)));
	CUF(dev_alloc(c, &dT, t2.size() * sizeof(double)));
	CUF(dev_alloc(c, &dh, 2 * (size_t)nBins * sizeof(unsigned long long)));
	CUF(dev_alloc(c, &scratch, rdf_scratch_bytes((int64_t)N, nullptr)));
	CUF(launch_download(c->g, c->T, c->prev, c->chains, nSlots, c->nSeeded, 0, flags, scan, tmp, tmpBytes, dx, dc, nullptr, nullptr, c->stream));
	CUF(cudaMemcpyAsync(dT, t2.data(), t2.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUF(cudaMemsetAsync(dh, 0, 2 * (size_t)nBins * sizeof(unsigned long long), c->stream));
	CUF(launch_rdf(c->g.L, dx, dc, (int64_t)N, dT, nBins, rMin, delta, lastEdge, scratch, dh, dh + nBins, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
	CUF(cudaMemcpyAsync(intra, dh, (size_t)nBins * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
	CUF(cudaMemcpyAsync(inter, dh + nBins, (size_t)nBins * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
#undef CUF
	cleanup();
	return 0;
}

int sarw_chain_shapes(sarw_ctx* c, double* ree2, double* rg2)
{
	if (!c || !ree2 || !rg2) return 1;
	if (!c->initiateCalled) return fail(c, "nothing to analyse before sarw_initiate");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	double* d = nullptr;
	const size_t n = c->g.nChains;
	CU(c, dev_alloc(c, &d, 2 * n * sizeof(double)));
	cudaError_t e = launch_chain_shapes(c->T, c->g.idAtomBase, c->g.nChains, (uint32_t)h.round, d, d + n, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(ree2, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(rg2, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
	dev_free(c, d);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return fail(c, "sarw_chain_shapes: %s", cudaGetErrorString(e));
	return 0;
}

int sarw_walk_checksum(sarw_ctx* c, uint64_t out[4])
{
	if (!c || !out) return 1;
	if (!c->initiateCalled) return fail(c, "nothing to sum before sarw_initiate");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	unsigned long long* d = nullptr;
	CU(c, dev_alloc(c, &d, 4 * sizeof(unsigned long long)));
	cudaError_t e = launch_walk_checksum(c->T, c->prev, c->g.idAtomBase, slots_for_rounds(c, c->hctl->round), d, c->numSMs, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	dev_free(c, d);
	if (e != cudaSuccess) return fail(c, "sarw_walk_checksum: %s", cudaGetErrorString(e));
	return 0;
}

int sarw_download_box(sarw_ctx* c, const double lo[3], const double hi[3], int64_t cap, int64_t* count, double* xyz, int64_t* ids, int64_t* prevIds, int32_t* chainID)
{
	if (!c || !lo || !hi || !count || cap < 0) return 1;
	if (cap > 0 && (!xyz || !ids || !prevIds || !chainID)) return fail(c, "sarw_download_box: null output buffer");
	if (!c->initiateCalled) return fail(c, "nothing to download before sarw_initiate");
	for (int k = 0; k < 3; k++) if (!(lo[k] >= 0 && lo[k] < hi[k] && hi[k] <= c->g.L[k])) return fail(c, "sarw_download_box: need 0 <= lo < hi <= boxSize in every dimension");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	const uint64_t nSlots = slots_for_rounds(c, c->hctl->round);
	const size_t n = (size_t)std::max<int64_t>(cap, 1);
	unsigned long long* dCount = nullptr; double* dx = nullptr; int64_t *di = nullptr, *dp = nullptr; int32_t* dc = nullptr;
	int rc = 0;
	auto cleanup = [&]() { dev_free(c, dCount); dev_free(c, dx); dev_free(c, di); dev_free(c, dp); dev_free(c, dc); };
#define CUF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
	CUF(dev_alloc(c, &dCount, sizeof(unsigned long long)));
	CUF(dev_alloc(c, &dx, n * 3 * sizeof(double)));
	CUF(dev_alloc(c, &di, n * sizeof(int64_t)));
	CUF(dev_alloc(c, &dp, n * sizeof(int64_t)));
	CUF(dev_alloc(c, &dc, n * sizeof(int32_t)));
	CUF(launch_gather_box(c->g, c->T, c->prev, nSlots, lo, hi, (uint64_t)cap, dCount, dx, di, dp, dc, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
	unsigned long long found = 0;
	CUF(cudaMemcpyAsync(&found, dCount, sizeof found, cudaMemcpyDeviceToHost, c->stream));
	CUF(cudaStreamSynchronize(c->stream));
	*count = (int64_t)found;
	const size_t m = (size_t)std::min<unsigned long long>(found, (unsigned long long)cap);
	if (m) {
		CUF(cudaMemcpyAsync(xyz, dx, m * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUF(cudaMemcpyAsync(ids, di, m * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
		CUF(cudaMemcpyAsync(prevIds, dp, m * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
		CUF(cudaMemcpyAsync(chainID, dc, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
		CUF(cudaStreamSynchronize(c->stream));
	}
#undef CUF
	cleanup();
	return 0;
}

// growth is over and no more overlap queries will be made: hand the cell table and the overflow table back (at 1e9 beads they are
// 100 GB that the analysis calls need for their own buffers). sarw_lookup_find_batch / growth calls fail afterwards.
int sarw_release_lookup(sarw_ctx* c)
{
	if (!c) return 1;
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->cellsVm) vm_free_cells(c); else dev_free(c, c->T.cells);
	dev_free(c, c->T.ovf);
	c->T.cells = nullptr; c->T.ovf = nullptr; c->lookupReleased = true;
	cudaStreamSynchronize(c->allocStream);
	return 0;
}

int sarw_chain_lengths(sarw_ctx* c, int32_t* lengths)
{
	if (!c || !lengths) return 1;
	cudaSetDevice(c->device);
	int32_t* d = nullptr;
	CU(c, dev_alloc(c, &d, (size_t)c->g.nChains * sizeof(int32_t)));
	cudaError_t e = launch_chain_lengths(c->chains, c->g.nChains, d, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(lengths, d, (size_t)c->g.nChains * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
	dev_free(c, d);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return fail(c, "sarw_chain_lengths: %s", cudaGetErrorString(e));
	return 0;
}

int sarw_round_counts(sarw_ctx* c, int64_t* counts, int64_t n)
{
	if (!c || (n > 0 && !counts)) return 1;
	if (!c->initiateCalled) return fail(c, "sarw_initiate must be called first");
	cudaSetDevice(c->device);
	if (fresh(c)) return 1;
	Control& h = *c->hctl;
	int64_t m = std::min<int64_t>(n, h.round);
	if (m <= 0) return 0;
	std::vector<uint32_t> acc((size_t)m);
	CU(c, cudaMemcpyAsync(acc.data(), c->roundAcc, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	int64_t run = 2ll * c->nSeeded;
	for (int64_t r = 0; r < m; r++) { run += acc[(size_t)r]; counts[r] = run; }
	return 0;
}

} // extern "C"
