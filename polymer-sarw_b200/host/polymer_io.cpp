#include "polymer_io.h"
#include "fast_format.h"
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <functional>
#include <thread>

namespace sarwhost {

void build_topology(Polymer& p)
{
	const int64_t N = p.nAtoms(), nC = p.nChains;
	p.angles.clear(); p.dihedrals.clear();
	if (N == 0 || nC == 0) return;
	// bucket atoms by chain, keeping acceptance order (counting sort: O(N))
	std::vector<int64_t> start((size_t)nC + 1, 0);
	for (int64_t i = 0; i < N; i++) start[(size_t)p.chainID[(size_t)i]]++;       // chainID is 1-based: counts land in [1..nC]
	for (int64_t c = 0; c < nC; c++) start[(size_t)c + 1] += start[(size_t)c];   // chain id c+1 now occupies [start[c], start[c+1])
	std::vector<int64_t> order((size_t)N), fill(start.begin(), start.end() - 1);
	for (int64_t i = 0; i < N; i++) order[(size_t)fill[(size_t)p.chainID[(size_t)i] - 1]++] = i;
	// per-chain output offsets (prefix sums), then the chains are written by several threads: every chain owns its slice of the output
	std::vector<int64_t> angAt((size_t)nC + 1, 0), dihAt((size_t)nC + 1, 0);
	for (int64_t c = 0; c < nC; c++) {
		const int64_t len = start[(size_t)c + 1] - start[(size_t)c];
		// sarw.cpp:406: the reference starts each chain at atom 2*iChain-2, i.e. it assumes the seeds of all chains exist
		const bool skip = len > 0 && order[(size_t)start[(size_t)c]] != 2 * c;
		angAt[(size_t)c + 1] = angAt[(size_t)c] + ((!skip && len >= 3) ? len - 2 : 0);
		dihAt[(size_t)c + 1] = dihAt[(size_t)c] + ((!skip && len >= 4) ? len - 3 : 0);
	}
	p.angles.resize((size_t)angAt[(size_t)nC] * 3); p.dihedrals.resize((size_t)dihAt[(size_t)nC] * 4);
	auto emit = [&](int64_t c0, int64_t c1) {
		for (int64_t c = c0; c < c1; c++) {
			const int64_t b = start[(size_t)c], len = start[(size_t)c + 1] - b;
			if (len > 0 && order[(size_t)b] != 2 * c) continue;
			int64_t ia = angAt[(size_t)c] * 3, id = dihAt[(size_t)c] * 4;
			for (int64_t k = 2; k < len; k++) {                                      // sarw.cpp:410
				p.angles[(size_t)ia++] = order[(size_t)(b + k - 2)] + 1; p.angles[(size_t)ia++] = order[(size_t)(b + k - 1)] + 1;
				p.angles[(size_t)ia++] = order[(size_t)(b + k)] + 1;
			}
			for (int64_t k = 3; k < len; k++)                                        // sarw.cpp:412
				for (int64_t q = 3; q >= 0; q--) p.dihedrals[(size_t)id++] = order[(size_t)(b + k - q)] + 1;
		}
	};
	int nT = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
	if (N < (1 << 20)) nT = 1;
	std::vector<std::thread> th;
	for (int t = 1; t < nT; t++) th.emplace_back(emit, nC * t / nT, nC * (t + 1) / nT);
	emit(0, nC / nT);
	for (auto& x : th) x.join();
}

namespace {

// format rows [0,n) with `row(i, buf)` (returns bytes written, < maxRow) on nThreads threads, write them in order; the next batch of
// chunks is formatted while the current one is being written (two sets of buffers)
bool write_rows(FILE* fp, int64_t n, size_t maxRow, int nThreads, const std::function<int(int64_t, char*)>& row)
{
	if (n <= 0) return true;
	if (nThreads <= 0) nThreads = (int)std::max(1u, std::thread::hardware_concurrency());
	const int64_t chunk = 1 << 16;
	const int64_t nChunks = (n + chunk - 1) / chunk;
	nThreads = (int)std::min<int64_t>(nThreads, nChunks);
	struct Batch { std::vector<std::vector<char>> bufs; std::vector<size_t> lens; int nb = 0; };
	Batch sets[2];
	for (Batch& s : sets) { s.bufs.resize((size_t)nThreads); s.lens.resize((size_t)nThreads); }
	auto format_batch = [&](Batch& s, int64_t c0) {
		s.nb = (int)std::min<int64_t>(nThreads, nChunks - c0);
		auto work = [&](int t) {
			const int64_t lo = (c0 + t) * chunk, hi = std::min(n, lo + chunk);
			std::vector<char>& b = s.bufs[(size_t)t];
			b.resize((size_t)(hi - lo) * maxRow);
			size_t at = 0;
			for (int64_t i = lo; i < hi; i++) at += (size_t)row(i, b.data() + at);
			s.lens[(size_t)t] = at;
		};
		std::vector<std::thread> th;
		for (int t = 1; t < s.nb; t++) th.emplace_back(work, t);
		work(0);
		for (auto& x : th) x.join();
	};
	int cur = 0;
	format_batch(sets[cur], 0);
	for (int64_t c0 = 0; c0 < nChunks; c0 += nThreads) {
		const int64_t next = c0 + nThreads;
		std::thread ahead;
		if (next < nChunks) ahead = std::thread(format_batch, std::ref(sets[cur ^ 1]), next);
		bool ok = true;
		for (int t = 0; t < sets[cur].nb && ok; t++)
			ok = fwrite(sets[cur].bufs[(size_t)t].data(), 1, sets[cur].lens[(size_t)t], fp) == sets[cur].lens[(size_t)t];
		if (ahead.joinable()) ahead.join();
		if (!ok) return false;
		cur ^= 1;
	}
	return true;
}

} // namespace

bool export_lammps(const Polymer& p, const std::string& path, int nThreads, std::string& error)
{
	FILE* fp = fopen(path.c_str(), "w");
	if (!fp) { error = "cannot open " + path + " for writing"; return false; }
	bool ok = true;
	// sarw.cpp:434-456; counts are %5d in the reference (int); %5lld prints the same text for N < 2^31 and keeps going beyond
	fprintf(fp, "%s chains\n\n", p.name.c_str());
	fprintf(fp, "%5lld\tatoms\n", (long long)p.nAtoms());
	fprintf(fp, "%5lld\tbonds\n", (long long)p.nBonds());
	fprintf(fp, "%5lld\tangles\n", (long long)p.nAngles());
	fprintf(fp, "%5lld\tdihedrals\n", (long long)p.nDihedrals());
	fprintf(fp, "\n");
	fprintf(fp, "2\tatom types\n1\tbond types\n1\tangle types\n1\tdihedral types\n");
	fprintf(fp, "\n");
	fprintf(fp, "0.0 %lf xlo xhi\n", p.box[0]);
	fprintf(fp, "0.0 %lf ylo yhi\n", p.box[1]);
	fprintf(fp, "0.0 %lf zlo zhi\n", p.box[2]);
	fprintf(fp, "\n");
	fprintf(fp, "Masses\n\n");
	fprintf(fp, "%d\t%lf\n", 1, 14.0 + 1);
	fprintf(fp, "%d\t%lf\n", 2, 14.0);
	fprintf(fp, "\n");
	fprintf(fp, "Atoms\n\n");                                                   // sarw.cpp:458-464
	ok = ok && write_rows(fp, p.nAtoms(), 160, nThreads, [&](int64_t i, char* b) {   // "%d\t%d\t%d\t%lf\t%lf\t%lf\n"
		char* q = put_int(b, (long long)(i + 1)); *q++ = '\t';
		q = put_int(q, p.chainID[(size_t)i]); *q++ = '\t';
		q = put_int(q, p.type[(size_t)i]);
		for (int k = 0; k < 3; k++) { *q++ = '\t'; q = put_fixed6(q, p.xyz[3 * (size_t)i + k]); }
		*q++ = '\n';
		return (int)(q - b); });
	fprintf(fp, "\n");
	fprintf(fp, "Bonds\n\n");                                                   // sarw.cpp:466-470
	ok = ok && write_rows(fp, p.nBonds(), 96, nThreads, [&](int64_t i, char* b) {      // "%d\t%d\t%d\t%d\n"
		char* q = put_int(b, (long long)(i + 1)); *q++ = '\t'; *q++ = '1';
		for (int k = 0; k < 2; k++) { *q++ = '\t'; q = put_int(q, (long long)p.bonds[2 * (size_t)i + k]); }
		*q++ = '\n';
		return (int)(q - b); });
	fprintf(fp, "\n");
	fprintf(fp, "Angles\n\n");                                                  // sarw.cpp:472-476
	ok = ok && write_rows(fp, p.nAngles(), 120, nThreads, [&](int64_t i, char* b) {
		char* q = put_int(b, (long long)(i + 1)); *q++ = '\t'; *q++ = '1';
		for (int k = 0; k < 3; k++) { *q++ = '\t'; q = put_int(q, (long long)p.angles[3 * (size_t)i + k]); }
		*q++ = '\n';
		return (int)(q - b); });
	fprintf(fp, "\n");
	fprintf(fp, "Dihedrals\n\n");                                               // sarw.cpp:478-485 (trailing tab on every row)
	ok = ok && write_rows(fp, p.nDihedrals(), 140, nThreads, [&](int64_t i, char* b) {
		char* q = put_int(b, (long long)(i + 1)); *q++ = '\t'; *q++ = '1';
		for (int k = 0; k < 4; k++) { *q++ = '\t'; q = put_int(q, (long long)p.dihedrals[4 * (size_t)i + k]); }
		*q++ = '\t'; *q++ = '\n';
		return (int)(q - b); });
	fprintf(fp, "\n");
	if (fclose(fp) != 0) ok = false;
	if (!ok) error = "write to " + path + " failed";
	return ok;
}

static int xyz_row(const Polymer& p, int64_t i, char* b)
{	char* q = b; *q++ = 'C';                                                   // "C\t%lf\t%lf\t%lf\n"
	for (int k = 0; k < 3; k++) { *q++ = '\t'; q = put_fixed6(q, p.xyz[3 * (size_t)i + k]); }
	*q++ = '\n';
	return (int)(q - b);
}

bool export_xyz(const Polymer& p, const std::string& path, int nThreads, std::string& error)
{
	FILE* fp = fopen(path.c_str(), "w");
	if (!fp) { error = "cannot open " + path + " for writing"; return false; }
	fprintf(fp, "%lld\n", (long long)p.nAtoms());                               // sarw.cpp:374-377
	fprintf(fp, "Polyethylene\n");
	bool ok = write_rows(fp, p.nAtoms(), 120, nThreads, [&](int64_t i, char* b) { return xyz_row(p, i, b); });
	if (fclose(fp) != 0) ok = false;
	if (!ok) error = "write to " + path + " failed";
	return ok;
}

bool export_xsf(const Polymer& p, const std::string& path, int nThreads, std::string& error)
{
	FILE* fp = fopen(path.c_str(), "w");
	if (!fp) { error = "cannot open " + path + " for writing"; return false; }
	fprintf(fp, "CRYSTAL\nPRIMVEC\n");                                          // sarw.cpp:385-393
	fprintf(fp, "%lf\t0.0\t0.0\n0.0\t%lf\t0.0\n0.0\t0.0\t%lf\n", p.box[0], p.box[1], p.box[2]);
	fprintf(fp, "CONVVEC\n");
	fprintf(fp, "%lf\t0.0\t0.0\n0.0\t%lf\t0.0\n0.0\t0.0\t%lf\n", p.box[0], p.box[1], p.box[2]);
	fprintf(fp, "PRIMCOORD\n%lld\t%d\n", (long long)p.nAtoms(), 1);
	bool ok = write_rows(fp, p.nAtoms(), 120, nThreads, [&](int64_t i, char* b) { return xyz_row(p, i, b); });
	if (fclose(fp) != 0) ok = false;
	if (!ok) error = "write to " + path + " failed";
	return ok;
}

bool export_chain_lengths(const Polymer& p, const std::string& path, std::string& error)
{
	FILE* fp = fopen(path.c_str(), "w");
	if (!fp) { error = "cannot open " + path + " for writing"; return false; }
	for (int32_t len : p.chainLengths) fprintf(fp, "%d\n", len);                 // sarw.cpp:507
	if (fclose(fp) != 0) { error = "write to " + path + " failed"; return false; }
	return true;
}

bool read_xyz_records(const std::string& path, std::vector<double>& out)
{
	std::ifstream f(path.c_str());
	if (!f.is_open()) return false;
	double x, y, z;
	while (f >> x >> y >> z) { out.push_back(x); out.push_back(y); out.push_back(z); }
	return true;
}

} // namespace sarwhost
