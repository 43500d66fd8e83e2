// Stand-in for JDFTx core/Util.h — see core/vector3.h in this directory for the why.
// Provides exactly what the reference touches: globalLog/logPrintf/die, mpiWorld{isHead,bcast},
// InitParams + initSystemCmdline (sarw.cpp:18,91-99; InputMap.cpp:30,38-40,51,62,81; config.in.h:8).
#ifndef STANDIN_CORE_UTIL_H
#define STANDIN_CORE_UTIL_H
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <core/string.h>

extern FILE* globalLog;
#define logPrintf(...) fprintf(globalLog, __VA_ARGS__)
#define logFlush() fflush(globalLog)
// die must be a braced block: InputMap.cpp:81 uses it without a trailing ';'
#define die(...) { fprintf(globalLog, __VA_ARGS__); if(globalLog != stdout) fprintf(stderr, __VA_ARGS__); exit(1); }

struct StandinMPI
{	bool isHead() const { return true; }
	void bcast(string&) const {}
};
extern StandinMPI* mpiWorld;

struct InitParams
{	const char* packageName;
	const char* versionString;
	const char* versionHash;
	const char* description;
	string inputFilename;
	string logFilename;
	bool dryRun;
	InitParams() : packageName(0), versionString(0), versionHash(0), description(0), dryRun(false) {}
};

//! parse -i <input> / -o <log>, open globalLog, print a banner
void initSystemCmdline(int argc, char** argv, InitParams& ip);

#endif
