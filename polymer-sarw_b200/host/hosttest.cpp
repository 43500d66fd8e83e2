// sarw_hosttest — exercises the host layer WITHOUT a GPU (CPU test-suite hook, tests/test_host.py):
//   sarw_hosttest parse <input>                 -> prints the resolved keys one per line
//   sarw_hosttest export <atoms.bin> <outdir>   -> topology + the four output files from a binary atom dump
// atoms.bin: int64 nChains, int64 N, int64 nBonds, double box[3], then xyz[3N] f64, chainID[N] i32, type[N] i32,
// bonds[2*nBonds] i64, chainLengths[nChains] i32; polymer name fixed by argv[4] (default Polyethylene).
#ifdef SARW_HOSTTEST
#include "input_map.h"
#include "polymer_io.h"
#include "fast_format.h"
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
using namespace sarwhost;

int main(int argc, char** argv)
{
	if (argc >= 3 && !strcmp(argv[1], "parse")) {
		InputMap m; std::string err;
		if (!m.load(argv[2], err)) { fputs(err.c_str(), stdout); return 1; }
		for (auto& kv : m.entries()) printf("%s=%s\n", kv.first.c_str(), kv.second.c_str());
		Vec3 b = m.getVector("boxSize", Vec3{ { 10, 10, 10 } });
		printf("#boxSize=%.17g,%.17g,%.17g nChains=%d minDist=%.17g maxTrials=%d\n", b[0], b[1], b[2], (int)m.get("nChains", 1), m.get("minDist", 2.0), (int)m.get("maxTrials", 100));
		return 0;
	}
	if (argc >= 4 && !strcmp(argv[1], "export")) {
		FILE* fp = fopen(argv[2], "rb");
		if (!fp) return 2;
		Polymer p; int64_t hdr[3];
		if (fread(hdr, sizeof(int64_t), 3, fp) != 3 || fread(p.box, sizeof(double), 3, fp) != 3) return 3;
		p.nChains = hdr[0]; const size_t N = (size_t)hdr[1], nB = (size_t)hdr[2];
		p.xyz.resize(3 * N); p.chainID.resize(N); p.type.resize(N); p.bonds.resize(2 * nB); p.chainLengths.resize((size_t)p.nChains);
		if (fread(p.xyz.data(), sizeof(double), 3 * N, fp) != 3 * N || fread(p.chainID.data(), 4, N, fp) != N || fread(p.type.data(), 4, N, fp) != N
			|| fread(p.bonds.data(), 8, 2 * nB, fp) != 2 * nB || fread(p.chainLengths.data(), 4, (size_t)p.nChains, fp) != (size_t)p.nChains) return 4;
		fclose(fp);
		p.name = argc >= 5 ? argv[4] : "Polyethylene";
		auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
		const double t0 = now();
		build_topology(p);
		const double t1 = now();
		std::string d = argv[3], err;
		const int threads = argc >= 6 ? atoi(argv[5]) : 0;
		if (!export_chain_lengths(p, d + "/chains.dat", err) || !export_lammps(p, d + "/polymer.dat", threads, err)) { fputs(err.c_str(), stderr); return 5; }
		const double t2 = now();
		if (!export_xyz(p, d + "/polymer.xyz", threads, err) || !export_xsf(p, d + "/polymer.xsf", threads, err)) { fputs(err.c_str(), stderr); return 5; }
		const double t3 = now();
		if (getenv("SARW_HOSTTEST_TIMING")) fprintf(stderr, "topology %.2f s, polymer.dat %.2f s, xyz + xsf %.2f s\n", t1 - t0, t2 - t1, t3 - t2);
		printf("%lld %lld %lld %lld\n", (long long)p.nAtoms(), (long long)p.nBonds(), (long long)p.nAngles(), (long long)p.nDihedrals());
		return 0;
	}
	if (argc >= 3 && !strcmp(argv[1], "fmtcheck")) {
		// put_fixed6 / put_int against snprintf on argv[2] pseudo-random doubles per family + ties + edge cases; prints the mismatch count
		const long n = atol(argv[2]);
		unsigned long long st = 0x9E3779B97F4A7C15ull, bad = 0, total = 0;
		auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
		auto check = [&](double x) {
			char a[64], b[64];
			*put_fixed6(a, x) = 0; snprintf(b, sizeof b, "%lf", x);
			total++;
			if (strcmp(a, b)) { if (bad < 5) fprintf(stderr, "mismatch %.17g: '%s' vs '%s'\n", x, a, b); bad++; }
		};
		const double edge[] = { 0.0, -0.0, 1e-7, -1e-7, 5e-7, 4.9999999999999996e-7, 5.0000000000000004e-7, 0.0078125, 0.0234375, 1.0, -1.0, 0.9999995, 0.99999949999999994,
			999999.9999995, 123456.7890125, 8589934591.9999990, 8589934592.0, 1e15, -1e30, 4.9e-324, 2.2250738585072014e-308, INFINITY, -INFINITY, NAN };
		for (double x : edge) check(x);
		for (long i = 0; i < n; i++) {
			const double u = (double)(rnd() >> 11) * (1.0 / 9007199254740992.0);
			check(u * 632.2); check(-u * 3000.0); check((u - 0.5) * 1e-4); check(std::ldexp((double)(rnd() >> 11), -(int)(rnd() % 80)));   // incl. many exact ties
			check((double)(long long)(rnd() % 20000000) / 128.0 + 1.0 / 256.0);                                                          // k/128 + 1/256: ties at 6 decimals
			char a[32], b[32]; const long long v = (long long)rnd() >> (rnd() % 63);
			*put_int(a, v) = 0; snprintf(b, sizeof b, "%lld", v); total++; if (strcmp(a, b)) bad++;
		}
		printf("%llu %llu\n", bad, total);
		return bad ? 1 : 0;
	}
	fprintf(stderr, "usage: sarw_hosttest parse <input> | export <atoms.bin> <outdir> [polymer] [threads]\n");
	return 64;
}
#endif
