// sarw — drop-in for the reference CLI `sarw -i <input> [-o <log>]` (README.md:22; sarw.cpp:13-89), with the chain-growth
// hot path running on a B200 through the C ABI of libsarw_b200.so (include/sarw_abi.h). No CPU fallback.
//
// Same input keys (union of what the code reads, sarw.cpp:22-38, and what README.md:26-41 documents), same log blocks
// (INPUTS / LOG FLAGS / PROGRESS / SUMMARY, sarw.cpp:41-56,64-77,491-503), same output files in the CWD
// (polymer.dat, polymer.xyz, polymer.xsf, chains.dat: sarw.cpp:373,383,433,506) plus the README's names
// (polymer.data, polymer.XYZ, polymer.XSF, README.md:44-46) as symlinks. Exit 0 after export even when jammed
// (sarw.cpp:88); exit 1 on a missing input file or required key (InputMap.cpp:30,51).
#include "input_map.h"
#include "polymer_io.h"
#include "sarw_abi.h"
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <unistd.h>
#include <sys/stat.h>
#include <vector>

using namespace sarwhost;

static FILE* globalLog = stdout;
static void logPrintf(const char* fmt, ...)
{	va_list ap; va_start(ap, fmt); vfprintf(globalLog, fmt, ap); va_end(ap);
}
[[noreturn]] static void die(const char* fmt, ...)
{	va_list ap; va_start(ap, fmt); vfprintf(globalLog, fmt, ap); va_end(ap);
	if (globalLog != stdout) { va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
	exit(1);
}
static double seconds_since(std::chrono::steady_clock::time_point t0)
{	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv)
{
	std::string inputFilename, logFilename, outDir;
	bool haveSeedOpt = false; unsigned long long seedOpt = 0; int device = -1, ioThreads = 0, gpus = 1, slabHalo = 0; bool timing = false;
	bool wantRdf = false; double rdfMin = 2.0, rdfDr = 0.5; int rdfBins = 20;   // scripts/rdf.py:84-85 defaults, 12 A range
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-i") && i + 1 < argc) inputFilename = argv[++i];
		else if (!strcmp(argv[i], "-o") && i + 1 < argc) logFilename = argv[++i];
		else if (!strcmp(argv[i], "--seed") && i + 1 < argc) { seedOpt = strtoull(argv[++i], 0, 10); haveSeedOpt = true; }
		else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--slab-halo") && i + 1 < argc) slabHalo = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--io-threads") && i + 1 < argc) ioThreads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--timing")) timing = true;
		else if (!strcmp(argv[i], "--rdf")) {   // optional "rMin,dr,nBins"
			wantRdf = true;
			if (i + 1 < argc && argv[i + 1][0] != '-') { if (sscanf(argv[++i], "%lf,%lf,%d", &rdfMin, &rdfDr, &rdfBins) != 3) { fprintf(stderr, "--rdf expects rMin,dr,nBins\n"); return 1; } }
		}
		else if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
			printf("usage: sarw -i <input> [-o <log>] [--seed N] [--device D] [--gpus N] [--io-threads T] [--timing] [--rdf [rMin,dr,nBins]]\n");
			return 0;
		}
	}
	if (!logFilename.empty()) { FILE* fp = fopen(logFilename.c_str(), "w"); if (fp) globalLog = fp; }
	logPrintf("\n*************** sarw (B200) 0.2.alpha ***************\nSelf Avoiding Random Walk\n");
	if (inputFilename.empty()) die("No input file given (use -i <file>).\n");

	InputMap inputMap; std::string err;
	if (!inputMap.load(inputFilename, err)) die("%s", err.c_str());

	// sarw.cpp:22-38, same defaults
	sarw_params P; memset(&P, 0, sizeof P);
	P.nChains = (int64_t)inputMap.get("nChains", 1);
	Vec3 boxSize = inputMap.getVector("boxSize", Vec3{ { 10, 10, 10 } });
	for (int k = 0; k < 3; k++) P.boxSize[k] = boxSize[k];
	P.minDist = inputMap.get("minDist", 2.0);
	P.targetMassDensity = inputMap.get("targetMassDensity", 0.92);
	bool haveNGraft = false;
	P.nGraftChains = (int64_t)inputMap.get("nGraftChains", 0, &haveNGraft);
	std::string graftLoc = inputMap.getString("graftLoc", "file");
	std::string growthOrder = inputMap.getString("growthOrder", "roundRobin");
	std::string boundary = inputMap.getString("boundary", "ppp");
	P.maxTrials = (int32_t)inputMap.get("maxTrials", 100);
	std::string polymer = inputMap.getString("polymer", "Polyethylene");
	bool logProgress = inputMap.getString("logProgress", "yes") == "yes";
	const bool logSteps = inputMap.getString("logSteps", "no") == "yes";
	if (logSteps) logProgress = true;                                           // sarw.cpp:511-516
	std::string obstacle = inputMap.getString("obstacle", "none");

	// README-only keys (README.md:32-33,40): not read by the reference code. graftFraction maps onto nGraftChains only
	// when it amounts to at least one chain and nGraftChains is absent (so test/pe.in, graftFraction 0.01 of 20 chains,
	// behaves exactly like the reference binary); roundRobin is accepted and, like in the reference, has no effect.
	std::vector<std::string> notes;
	bool haveFrac = false; double graftFraction = inputMap.get("graftFraction", 0.0, &haveFrac);
	if (haveFrac && !haveNGraft) {
		int64_t n = (int64_t)nearbyint(graftFraction * (double)P.nChains);
		if (n > 0) { P.nGraftChains = n; notes.push_back("graftFraction -> nGraftChains = " + std::to_string(n) + " (graftLoc " + graftLoc + ")"); }
	}
	bool haveBias = false; Vec3 bias = inputMap.getVector("growthBias", Vec3{ { 0, 0, 0 } }, &haveBias);
	for (int k = 0; k < 3; k++) P.growthBias[k] = bias[k];
	if (haveBias && (bias[0] != 0 || bias[1] != 0 || bias[2] != 0))
		notes.push_back("growthBias: trials whose step has growthBias.step < 0 are rejected (no reference implementation exists)");
	if (inputMap.has("roundRobin") && inputMap.getString("roundRobin", "yes") != "yes")
		notes.push_back("roundRobin no: the reference code ignores this key and grows round-robin; so does this build");
	bool haveSeedKey = false; double seedKey = inputMap.get("seed", 0, &haveSeedKey);
	if (haveSeedOpt) P.seed = seedOpt;
	else if (haveSeedKey) P.seed = (uint64_t)seedKey;
	else P.seed = (uint64_t)time(NULL) * 2654435761ull + (uint64_t)getpid();    // sarw.cpp:16: a different walk every run
	P.device = device;

	P.graftLoc = graftLoc == "file" ? SARW_GRAFT_FILE : graftLoc == "z" ? SARW_GRAFT_Z : graftLoc == "zCenter" ? SARW_GRAFT_ZCENTER : SARW_GRAFT_OTHER;
	P.growthOrder = growthOrder == "grafted" ? SARW_ORDER_GRAFTED : SARW_ORDER_ROUNDROBIN;
	P.boundary = boundary == "ppf" ? SARW_BOUNDARY_PPF : boundary == "pfp" ? SARW_BOUNDARY_PFP : boundary == "fpp" ? SARW_BOUNDARY_FPP : SARW_BOUNDARY_PPP;

	// sarw.cpp:41-56
	logPrintf("\nINPUTS:\n");
	logPrintf("nChains = %d\n", (int)P.nChains);
	logPrintf("boxSize = (%lg x %lg x %lg)\n", P.boxSize[0], P.boxSize[1], P.boxSize[2]);
	logPrintf("minDist = %lg\n", P.minDist);
	logPrintf("targetMassDensity = %lg\n", P.targetMassDensity);
	logPrintf("nGraftChains = %d\n", (int)P.nGraftChains);
	logPrintf("graftLoc = %s\n", graftLoc.c_str());
	logPrintf("growthOrder = %s\n", growthOrder.c_str());
	logPrintf("boundary = %s\n", boundary.c_str());
	logPrintf("maxTrials = %d\n", P.maxTrials);
	logPrintf("polymer = %s\n", polymer.c_str());
	logPrintf("obstacle = %s\n", obstacle.c_str());
	logPrintf("\nLOG FLAGS:\n");
	logPrintf("logProgress = %s\n", logProgress ? "yes" : "no");
	logPrintf("logSteps = %s\n", logSteps ? "yes" : "no");
	if (!notes.empty() || haveSeedKey || haveSeedOpt) {
		logPrintf("\nB200 NOTES:\n");
		if (haveSeedKey || haveSeedOpt) logPrintf("seed = %llu\n", (unsigned long long)P.seed);
		for (const std::string& n : notes) logPrintf("%s\n", n.c_str());
		if (boundary != "ppp" && boundary != "ppf" && boundary != "pfp" && boundary != "fpp")
			logPrintf("boundary %s: only ppf, pfp, fpp act in the reference (sarw.cpp:201-203); treated as ppp\n", boundary.c_str());
	}

	auto tAll = std::chrono::steady_clock::now();
	// --gpus N: ONE melt over N GPUs. The box is decomposed into N slabs along z, one context (rank) per slab on device (D + r) mod #devices;
	// every rank gets the same parameters, grafts and obstacles; the complete walk is downloaded from rank 0.
	if (gpus < 1 || gpus > 8) die("--gpus must be between 1 and 8\n");
	std::vector<sarw_ctx*> ranks((size_t)gpus, nullptr);
	for (int r = 0; r < gpus; r++) {
		sarw_params Pr = P;
		if (gpus > 1) {
			const int nDev = sarw_device_count();
			if (nDev < 1) die("\nno CUDA device available; sarw (B200) has no CPU fallback\n");
			Pr.device = ((device < 0 ? 0 : device) + r) % nDev;
			Pr.slabRank = r; Pr.slabRanks = gpus; Pr.slabHalo = slabHalo; Pr.launchMode = 1;
		}
		if (sarw_create(&Pr, &ranks[(size_t)r]) != 0) die("\nsarw_create failed: %s\n", sarw_last_error(nullptr));
	}
	sarw_ctx* ctx = ranks[0];
	if (gpus > 1) notes.clear(), logPrintf("\nB200: one melt over %d GPUs, box decomposed into %d slabs along z (%.1f MB of cell list per rank)\n", gpus, gpus, (double)sarw_cell_table_bytes(ctx) / 1e6);

	// sarw.cpp:59, 102-123
	if (P.nGraftChains > 0 && graftLoc == "file") {
		std::vector<double> grafts;
		if (read_xyz_records("grafts.dat", grafts)) {
			int64_t n = (int64_t)grafts.size() / 3;
			logPrintf("\nReading grafts.dat: %d graft seeds found\n", (int)n);
			if (n < P.nGraftChains) die("Inadequate number of seeds! Expected %d seeds.\n", (int)P.nGraftChains);   // the reference only logs, then reads out of bounds
			if (n > P.nGraftChains) { logPrintf("Removing extra seeds!\n"); n = P.nGraftChains; }
			for (sarw_ctx* c : ranks) if (sarw_set_grafts(c, grafts.data(), n) != 0) die("%s\n", sarw_last_error(c));
		} else die("\nFailed to open grafts.dat, needed for %d grafted chains.\n", (int)P.nGraftChains);
	}
	// sarw.cpp:62, 126-145
	if (obstacle != "none") {
		std::vector<double> obst;
		if (read_xyz_records(obstacle, obst)) {
			logPrintf("\nReading %s: %d obstacles found\n", obstacle.c_str(), (int)(obst.size() / 3));
			for (sarw_ctx* c : ranks) if (sarw_add_obstacles(c, obst.data(), (int64_t)obst.size() / 3) != 0) die("%s\n", sarw_last_error(c));
		} else logPrintf("\nFailed to open %s. Assuming no obstacles.\n", obstacle.c_str());
	}

	// sarw.cpp:64-80: the loop runs on the device; the PROGRESS line is reproduced from the per-round bead counts
	if (logProgress) logPrintf("\nPROGRESS: ");
	auto tGrow = std::chrono::steady_clock::now();
	int32_t ok = 0;
	sarw_stats st; memset(&st, 0, sizeof st);
	for (sarw_ctx* c : ranks) if (sarw_initiate(c, &ok) != 0) die("\n%s\n", sarw_last_error(c));
	if (gpus == 1) { if (sarw_grow(ctx, 0, -1, &st) != 0) die("\n%s\n", sarw_last_error(ctx)); }
	else if (sarw_mr_grow_local(ranks.data(), gpus, -1, &st) != 0) die("\n%s\n", sarw_last_error(ctx));
	if (sarw_terminate(ctx) != 0) die("\n%s\n", sarw_last_error(ctx));
	const double growSeconds = seconds_since(tGrow);
	if (logProgress) {
		std::vector<int64_t> counts((size_t)st.rounds);
		if (st.rounds > 0 && sarw_round_counts(ctx, counts.data(), st.rounds) != 0) die("\n%s\n", sarw_last_error(ctx));
		int oldProgress = 0;
		for (int64_t r = 0; r < st.rounds; r++) {
			int currentProgress = (int)(100. * (double)counts[(size_t)r] / (double)st.target);   // sarw.cpp:70
			if (currentProgress > oldProgress) { oldProgress = currentProgress; logPrintf("%d, ", currentProgress); }
		}
		logPrintf("\n");
	}

	Polymer pm;
	pm.name = polymer; pm.nChains = P.nChains;
	for (int k = 0; k < 3; k++) pm.box[k] = P.boxSize[k];
	pm.xyz.resize((size_t)st.nBeads * 3); pm.chainID.resize((size_t)st.nBeads); pm.type.resize((size_t)st.nBeads);
	pm.bonds.resize((size_t)st.nBonds * 2); pm.chainLengths.resize((size_t)P.nChains);
	auto tDown = std::chrono::steady_clock::now();
	if (sarw_download(ctx, pm.xyz.data(), pm.chainID.data(), pm.type.data(), pm.bonds.data()) != 0) die("\n%s\n", sarw_last_error(ctx));
	if (sarw_chain_lengths(ctx, pm.chainLengths.data()) != 0) die("\n%s\n", sarw_last_error(ctx));
	const double downSeconds = seconds_since(tDown);
	sarw_get_stats(ctx, &st);
	double rdfSeconds = 0;
	if (wantRdf) {
		// what scripts/rdf.py computes off line, on the device: pair histograms (rdf.py:13-20,39-57) + chain shapes
		auto tRdf = std::chrono::steady_clock::now();
		std::vector<uint64_t> intra((size_t)rdfBins), inter((size_t)rdfBins);
		std::vector<double> ree2((size_t)P.nChains), rg2((size_t)P.nChains);
		if (sarw_rdf(ctx, rdfMin, rdfDr, rdfBins, intra.data(), inter.data()) != 0) die("\n%s\n", sarw_last_error(ctx));
		if (sarw_chain_shapes(ctx, ree2.data(), rg2.data()) != 0) die("\n%s\n", sarw_last_error(ctx));
		rdfSeconds = seconds_since(tRdf);
		FILE* fp = fopen("rdf.dat", "w");
		if (!fp) die("Failed to open rdf.dat for writing\n");
		const double N = (double)st.nBeads, V = P.boxSize[0] * P.boxSize[1] * P.boxSize[2];
		fprintf(fp, "# r_lo r_hi intra_pairs inter_pairs g_intra g_inter   (ordered pairs as scripts/rdf.py counts them; g = pairs / (N(N-1)/V * shell volume))\n");
		for (int k = 0; k < rdfBins; k++) {
			const double lo = rdfMin + k * rdfDr, hi = rdfMin + (k + 1) * rdfDr;
			const double ideal = N * (N - 1) / V * 4.0 / 3.0 * M_PI * (hi * hi * hi - lo * lo * lo);
			fprintf(fp, "%.6f %.6f %llu %llu %.6e %.6e\n", lo, hi, (unsigned long long)intra[(size_t)k], (unsigned long long)inter[(size_t)k],
				ideal > 0 ? (double)intra[(size_t)k] / ideal : 0.0, ideal > 0 ? (double)inter[(size_t)k] / ideal : 0.0);
		}
		fclose(fp);
		fp = fopen("shapes.dat", "w");
		if (!fp) die("Failed to open shapes.dat for writing\n");
		fprintf(fp, "# chain length Ree^2 Rg^2\n");
		for (int64_t c = 0; c < P.nChains; c++) fprintf(fp, "%lld %d %.10e %.10e\n", (long long)c + 1, pm.chainLengths[(size_t)c], ree2[(size_t)c], rg2[(size_t)c]);
		fclose(fp);
		logPrintf("\nExporting pair distribution and chain shapes: rdf.dat, shapes.dat\n");
	}
	for (sarw_ctx* c : ranks) sarw_destroy(c);

	// sarw.cpp:489-509
	const long long vol = (long long)(P.boxSize[0] * P.boxSize[1] * P.boxSize[2]);
	const double actualMassDensity = 14.0 * (double)st.nBeads / (6.022e23 * (double)vol * 1e-24);
	logPrintf("\nSUMMARY:\n");
	logPrintf("No. of united atoms: ");
	logPrintf("desired = %lld, ", (long long)st.target);
	logPrintf("actual = %lld\n", (long long)st.nBeads);
	logPrintf("Number density: ");
	logPrintf("desired = %.3e, ", 6.022e23 * P.targetMassDensity / 14.0);
	logPrintf("actual = %.3e\n", 6.022e23 * actualMassDensity / 14.0);
	logPrintf("Mass density: ");
	logPrintf("desired = %.2f, ", P.targetMassDensity);
	logPrintf("actual = %.2f\n", actualMassDensity);
	logPrintf("\nExporting distribution of chain lengths: chains.dat\n");
	if (!export_chain_lengths(pm, "chains.dat", err)) die("%s\n", err.c_str());

	auto tTopo = std::chrono::steady_clock::now();
	build_topology(pm);
	const double topoSeconds = seconds_since(tTopo);
	auto tExp = std::chrono::steady_clock::now();
	logPrintf("\nExporting LAMMPS data file: polymer.dat\n");
	if (!export_lammps(pm, "polymer.dat", ioThreads, err)) die("%s\n", err.c_str());
	logPrintf("Exporting XYZ file: polymer.xyz\n");
	if (!export_xyz(pm, "polymer.xyz", ioThreads, err)) die("%s\n", err.c_str());
	logPrintf("Exporting XSF file: polymer.xsf\n");
	if (!export_xsf(pm, "polymer.xsf", ioThreads, err)) die("%s\n", err.c_str());
	const double expSeconds = seconds_since(tExp);
	// README.md:44-46 names, as aliases
	const char* alias[3][2] = { { "polymer.dat", "polymer.data" }, { "polymer.xyz", "polymer.XYZ" }, { "polymer.xsf", "polymer.XSF" } };
	for (auto& a : alias) {
		struct stat sb;   // only an old alias is replaced: on a case-insensitive file system polymer.XYZ IS polymer.xyz, the file just written
		if (lstat(a[1], &sb) == 0) { if (!S_ISLNK(sb.st_mode)) continue; unlink(a[1]); }
		if (symlink(a[0], a[1]) != 0) { /* read-only or link-less file systems: skip */ }
	}

	if (timing) {
		logPrintf("\nTIMING (B200):\n");
		logPrintf("seeding + growth: %.3f ms on device (%.3f s wall incl. launches), %lld rounds, %.3e accepted beads/s\n",
			st.seedMs + st.growMs, growSeconds, (long long)st.rounds, (double)st.nBeads / ((st.seedMs + st.growMs) * 1e-3));
		logPrintf("trials per accepted bead (reference terms): %.2f; exact overlap queries evaluated: %llu; same-round conflicts: %llu; dead chains: %lld\n",
			(double)st.trials / (double)(st.nBeads ? st.nBeads : 1), (unsigned long long)st.queries, (unsigned long long)st.conflicts, (long long)st.deadChains);
		logPrintf("download %.3f s, topology %.3f s, export %.3f s, total %.3f s\n", downSeconds, topoSeconds, expSeconds, seconds_since(tAll));
		if (wantRdf) logPrintf("analysis (rdf + chain shapes) %.3f s\n", rdfSeconds);
	}
	if (globalLog != stdout) fclose(globalLog);
	return 0;
}
