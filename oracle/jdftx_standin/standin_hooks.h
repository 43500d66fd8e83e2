// Hooks the harness uses to swap the stand-in RNG (see core/Random.h in this directory).
#ifndef STANDIN_HOOKS_H
#define STANDIN_HOOKS_H
extern "C" {
void standin_rng_set_hooks(void (*seedFn)(int), double (*uniform01Fn)());
unsigned long long standin_rng_seed_calls();
unsigned long long standin_rng_uniform_calls();
void standin_rng_reset_counters();
}
#endif
