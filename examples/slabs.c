/* slabs.c — ONE melt over several GPUs from plain C: the box is decomposed into N slabs along z, one context per slab.
 *   gcc -std=c99 -Iinclude examples/slabs.c -Lpolymer-sarw_b200/lib -lsarw_b200 -Wl,-rpath,$PWD/polymer-sarw_b200/lib -o slabs
 *   ./slabs [N]        (default 2; devices are used round-robin, so N slabs also run on fewer GPUs)
 * No counterpart in the reference (it is serial and its lookup table is one array, PeriodicLookup.h:33); this is what `sarw --gpus N` does. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sarw_abi.h"

int main(int argc, char** argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 2;
	sarw_ctx* ctx[8] = { 0 };
	sarw_params P;
	sarw_stats st;
	int32_t ok = 0, layers[4];
	int r, nDev = sarw_device_count();
	if (n < 1 || n > 8) { fprintf(stderr, "1..8 slabs\n"); return 1; }
	if (nDev < 1) { fprintf(stderr, "no CUDA device: libsarw_b200 has no CPU fallback\n"); return 1; }
	memset(&P, 0, sizeof P);
	P.nChains = 400; P.boxSize[0] = P.boxSize[1] = P.boxSize[2] = 60.0;              /* 400 chains, ~20 beads each */
	P.minDist = 1.0; P.targetMassDensity = 0.92; P.maxTrials = 300; P.seed = 1; P.launchMode = 1;
	for (r = 0; r < n; r++) {
		P.device = r % nDev;
		if (n > 1) { P.slabRank = r; P.slabRanks = n; }                                /* this context holds slab r of n (+ halo) of the cell list */
		if (sarw_create(&P, &ctx[r])) { fprintf(stderr, "%s\n", sarw_last_error(NULL)); return 1; }
		if (sarw_initiate(ctx[r], &ok)) { fprintf(stderr, "%s\n", sarw_last_error(ctx[r])); return 1; }   /* every rank seeds all chains */
	}
	if (sarw_mr_grow_local(ctx, n, -1, &st)) { fprintf(stderr, "%s\n", sarw_last_error(ctx[0])); return 1; }
	if (sarw_terminate(ctx[0])) return 1;
	printf("%lld beads in %lld rounds on %d slab(s); %lld bytes of cell list on rank 0\n", (long long)st.nBeads, (long long)st.rounds, n,
		(long long)sarw_cell_table_bytes(ctx[0]));
	if (n > 1 && sarw_slab_layers(29, n, 0, 8, layers) == 0)                           /* the decomposition itself is a pure function */
		printf("29 layers over %d ranks: rank 0 owns [%d, %d), holds %d layers from layer %d\n", n, layers[0], layers[1], layers[3], layers[2]);
	{	/* every rank holds the complete walk: download from any of them */
		uint64_t sum0[4], sumr[4];
		if (sarw_walk_checksum(ctx[0], sum0) || sarw_walk_checksum(ctx[n - 1], sumr)) return 1;
		printf("walk checksum on rank 0 and rank %d %s\n", n - 1, memcmp(sum0, sumr, sizeof sum0) == 0 ? "agree" : "DIFFER");
	}
	for (r = 0; r < n; r++) sarw_destroy(ctx[r]);
	return 0;
}
