/* mc33_testgrids.c -- TEST INFRASTRUCTURE ONLY.
 * Deterministic synthetic grids used by the known-answer tests (BASELINE.md
 * section 2, K3..K6) and by bench.py's cpu_baseline leg.  x-fastest fill order. */
#include <stdint.h>
#include <stddef.h>

static inline double xs64(uint64_t *s)
{
	uint64_t v = *s;
	v ^= v << 13; v ^= v >> 7; v ^= v << 17;
	*s = v;
	return (double)(v >> 11) * (1.0 / 9007199254740992.0); /* 2^-53 */
}

/* kind 0: float (float)(2r-1); 1: double 2r-1; 2: u8 (uint8)(r*scale); 3: u16 (uint16)(r*scale) */
void mc33o_fill_xorshift(int kind, void *dst, uint64_t n, uint64_t seed, double scale)
{
	uint64_t s = seed ? seed : 88172645463325252ull;
	for (uint64_t i = 0; i < n; i++) {
		double r = xs64(&s);
		switch (kind) {
		case 0: ((float *)dst)[i] = (float)(2 * r - 1); break;
		case 1: ((double *)dst)[i] = 2 * r - 1; break;
		case 2: ((uint8_t *)dst)[i] = (uint8_t)(r * scale); break;
		default: ((uint16_t *)dst)[i] = (uint16_t)(r * scale); break;
		}
	}
}
