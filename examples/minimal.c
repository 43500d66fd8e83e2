/* minimal.c — the C ABI of libsarw_b200 from plain C: grow one melt, print its size, analyse it on the device.
 *   gcc -std=c99 -Iinclude examples/minimal.c -Lpolymer-sarw_b200/lib -lsarw_b200 -Wl,-rpath,$PWD/polymer-sarw_b200/lib -o minimal
 * Replaces what main() of the reference does between reading the input and exporting (sarw.cpp:22-79). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sarw_abi.h"

int main(void)
{
	sarw_params P;
	sarw_ctx* ctx = NULL;
	sarw_stats st;
	int32_t ok = 0;
	memset(&P, 0, sizeof P);
	P.nChains = 20; P.boxSize[0] = 10; P.boxSize[1] = 10; P.boxSize[2] = 40;        /* test/pe.in */
	P.minDist = 1.0; P.targetMassDensity = 0.92; P.maxTrials = 1000;
	P.graftLoc = SARW_GRAFT_FILE; P.growthOrder = SARW_ORDER_ROUNDROBIN; P.boundary = SARW_BOUNDARY_PPF;
	P.seed = 1; P.device = -1;
	if (sarw_create(&P, &ctx)) { fprintf(stderr, "%s\n", sarw_last_error(NULL)); return 1; }
	if (sarw_initiate(ctx, &ok) || sarw_grow(ctx, 0, -1, &st) || sarw_terminate(ctx)) { fprintf(stderr, "%s\n", sarw_last_error(ctx)); return 1; }
	printf("%lld beads, %lld bonds in %lld rounds (target %lld); %.3f ms on the device\n", (long long)st.nBeads, (long long)st.nBonds,
		(long long)st.rounds, (long long)st.target, st.seedMs + st.growMs);
	{
		double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)st.nBeads);
		int32_t* chain = (int32_t*)malloc(sizeof(int32_t) * (size_t)st.nBeads);
		uint64_t intra[20], inter[20];
		if (sarw_download(ctx, xyz, chain, NULL, NULL) || sarw_rdf(ctx, 2.0, 0.5, 20, intra, inter)) { fprintf(stderr, "%s\n", sarw_last_error(ctx)); return 1; }
		printf("first bead of chain %d at (%.6f, %.6f, %.6f); %llu intra-chain pairs in [2.5, 3.0) A\n", chain[0], xyz[0], xyz[1], xyz[2],
			(unsigned long long)intra[1]);
		free(xyz); free(chain);
	}
	sarw_destroy(ctx);
	return 0;
}
