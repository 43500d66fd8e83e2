// fast_format.h — the two conversions the exporters spend their time in, without printf:
//   put_int(p, v)      decimal of a signed integer                         == "%lld"
//   put_fixed6(p, x)   fixed notation with six decimals                     == "%lf" (glibc: the exact binary value, round-half-even)
// Both return the advanced pointer. put_fixed6 is exact, not approximately right: |x| = m * 2^e with a 53-bit m, so
// |x| * 10^6 = (m * 10^6) * 2^e is formed in 128-bit integer arithmetic and rounded half-to-even on the discarded bits, which is what
// printf does with the exact decimal expansion. Values it does not cover (|x| >= 2^33, inf, nan) go through snprintf.
// tests/test_host.py::test_fast_format_equals_printf compares it with snprintf on random, tie and edge-case inputs.
#ifndef SARW_FAST_FORMAT_H
#define SARW_FAST_FORMAT_H
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace sarwhost {

inline char* put_uint(char* p, unsigned long long v)
{
	char tmp[24]; int n = 0;
	do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
	while (n) *p++ = tmp[--n];
	return p;
}
inline char* put_int(char* p, long long v)
{
	if (v < 0) { *p++ = '-'; return put_uint(p, 0ull - (unsigned long long)v); }
	return put_uint(p, (unsigned long long)v);
}
inline char* put_fixed6(char* p, double x)
{
	uint64_t bits; memcpy(&bits, &x, sizeof bits);
	const bool neg = (bits >> 63) != 0;
	const int expField = (int)((bits >> 52) & 0x7FF);
	const uint64_t frac = bits & ((1ull << 52) - 1);
	if (expField == 0x7FF || expField >= 1023 + 33) {                    // inf / nan / |x| >= 2^33 (never a coordinate): printf, at most 47 characters
		const int n = snprintf(p, 48, "%lf", x);
		return p + (n > 47 ? 47 : n);
	}
	const uint64_t m = expField ? (frac | (1ull << 52)) : frac;          // |x| = m * 2^e
	const int e = (expField ? expField : 1) - 1075;                      // -1074 .. -1043+...; here e <= 33 - 53 < 0
	unsigned __int128 prod = (unsigned __int128)m * 1000000u;           // < 2^73
	unsigned long long q;
	const int s = -e;                                                    // bits to discard, >= 20
	if (s >= 128) q = 0;                                                 // |x| * 10^6 < 2^(73-128): rounds to zero
	else {
		const unsigned __int128 one = 1;
		const unsigned __int128 kept = prod >> s, rem = prod & ((one << s) - 1), half = one << (s - 1);
		q = (unsigned long long)kept;
		if (rem > half || (rem == half && (q & 1ull))) ++q;
	}
	if (neg) *p++ = '-';
	p = put_uint(p, q / 1000000ull);
	*p++ = '.';
	unsigned f = (unsigned)(q % 1000000ull);
	for (int k = 5; k >= 0; --k) { p[k] = (char)('0' + f % 10); f /= 10; }
	return p + 6;
}

} // namespace sarwhost
#endif
