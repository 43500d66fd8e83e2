// polymer_io.h — what happens after the hot path: topology (sarw.cpp:397-422) and the exporters / report
// (sarw.cpp:370-395, 427-509), byte-compatible with the reference, O(N) and multi-threaded.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sarwhost {

// the public state main()'s exporters read (sarw.h:77-83), as sarw_download returns it
struct Polymer {
	std::string name;                 // SARW::polymer, header of polymer.dat
	double box[3];
	int64_t nChains = 0;
	std::vector<double> xyz;          // 3*N, acceptance order (never wrapped into the box, sarw.cpp:184)
	std::vector<int32_t> chainID;     // 1-based
	std::vector<int32_t> type;        // 1 = CH3, 2 = CH2
	std::vector<int64_t> bonds;       // 2*nBonds, 1-based atom ids, type always 1 (sarw.cpp:272,309)
	std::vector<int32_t> chainLengths;
	std::vector<int64_t> angles;      // 3*nAngles, filled by build_topology
	std::vector<int64_t> dihedrals;   // 4*nDihedrals
	int64_t nAtoms() const { return (int64_t)chainID.size(); }
	int64_t nBonds() const { return (int64_t)bonds.size() / 2; }
	int64_t nAngles() const { return (int64_t)angles.size() / 3; }
	int64_t nDihedrals() const { return (int64_t)dihedrals.size() / 4; }
};

// SARW::calcAnglesDihedrals (sarw.cpp:397-422) in O(N): consecutive triples / quads along each chain, chains in id
// order. The reference scans all atoms once per chain (O(nChains*N): 446 s at 1e7 beads).
void build_topology(Polymer& p);

// exportLAMMPS / exportXYZ / exportXSF / chains.dat (sarw.cpp:427-487, 370-379, 381-395, 505-508); false on I/O error
// (the reference leaves fopen unchecked). nThreads <= 0: hardware concurrency.
bool export_lammps(const Polymer& p, const std::string& path, int nThreads, std::string& error);
bool export_xyz(const Polymer& p, const std::string& path, int nThreads, std::string& error);
bool export_xsf(const Polymer& p, const std::string& path, int nThreads, std::string& error);
bool export_chain_lengths(const Polymer& p, const std::string& path, std::string& error);

// graft / obstacle files: whitespace-separated "x y z" records until the first parse failure (sarw.cpp:110,135)
bool read_xyz_records(const std::string& path, std::vector<double>& out);

} // namespace sarwhost
