/* oracle/xzo_tables.c -- see xzo_tables.h (TEST INFRASTRUCTURE ONLY). */
#include <string.h>
#include "xzo.h"
#include "xzo_tables.h"

uint32_t xzo_crc32_table[256];
uint64_t xzo_crc64_table[256];
uint8_t xzo_rc_prices[128];
static int g_init;

void xzo_tables_init(void)
{
	if (g_init) return;
	for (uint32_t b = 0; b < 256; ++b) {
		uint32_t r = b;
		for (int i = 0; i < 8; ++i) r = (r & 1) ? (r >> 1) ^ 0xEDB88320u : r >> 1;
		xzo_crc32_table[b] = r;
		uint64_t q = b;
		for (int i = 0; i < 8; ++i) q = (q & 1) ? (q >> 1) ^ 0xC96C5795D7870F42ull : q >> 1;
		xzo_crc64_table[b] = q;
	}
	/* price_tablegen.c:28-56 */
	for (uint32_t i = 8; i < 2048; i += 16) {
		uint32_t w = i, bit_count = 0;
		for (int j = 0; j < 4; ++j) {
			w *= w; bit_count <<= 1;
			while (w >= (1u << 16)) { w >>= 1; ++bit_count; }
		}
		xzo_rc_prices[i >> 4] = (uint8_t)((11 << 4) - 15 - bit_count);
	}
	g_init = 1;
}

/* lzma_crc32, check/crc32_fast.c (bytewise form of the same polynomial division) */
uint32_t xzo_crc32(const uint8_t *buf, size_t size, uint32_t crc)
{
	xzo_tables_init();
	crc = ~crc;
	while (size--) crc = xzo_crc32_table[(crc ^ *buf++) & 0xFF] ^ (crc >> 8);
	return ~crc;
}

/* lzma_crc64, check/crc64_fast.c:48-92 */
uint64_t xzo_crc64(const uint8_t *buf, size_t size, uint64_t crc)
{
	xzo_tables_init();
	crc = ~crc;
	while (size--) crc = xzo_crc64_table[(crc ^ *buf++) & 0xFF] ^ (crc >> 8);
	return ~crc;
}


/* SHA-256 (FIPS 180-4) for LZMA_CHECK_SHA256; the reference's own implementation is check/sha256.c.
 * Straightforward form with the full 64-word message schedule. */
static uint32_t ror32(uint32_t x, unsigned n) { return (x >> n) | (x << (32 - n)); }
void xzo_sha256(const uint8_t *buf, size_t size, uint8_t out[32])
{
	static const uint32_t K[64] = {
		0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
		0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
		0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
		0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
		0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
		0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2 };
	uint32_t H[8] = { 0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19 };
	const uint64_t bits = (uint64_t)size * 8;
	const size_t padded = ((size + 8) / 64 + 1) * 64;
	for (size_t off = 0; off < padded; off += 64) {
		uint8_t blk[64];
		for (size_t i = 0; i < 64; ++i) {
			const size_t p = off + i;
			if (p < size) blk[i] = buf[p];
			else if (p == size) blk[i] = 0x80;
			else if (p >= padded - 8) blk[i] = (uint8_t)(bits >> (8 * (padded - 1 - p)));
			else blk[i] = 0;
		}
		uint32_t W[64];
		for (int t = 0; t < 16; ++t) W[t] = ((uint32_t)blk[4 * t] << 24) | ((uint32_t)blk[4 * t + 1] << 16) | ((uint32_t)blk[4 * t + 2] << 8) | blk[4 * t + 3];
		for (int t = 16; t < 64; ++t) {
			const uint32_t s0 = ror32(W[t - 15], 7) ^ ror32(W[t - 15], 18) ^ (W[t - 15] >> 3);
			const uint32_t s1 = ror32(W[t - 2], 17) ^ ror32(W[t - 2], 19) ^ (W[t - 2] >> 10);
			W[t] = W[t - 16] + s0 + W[t - 7] + s1;
		}
		uint32_t v[8];
		memcpy(v, H, sizeof(v));
		for (int t = 0; t < 64; ++t) {
			const uint32_t T1 = v[7] + (ror32(v[4], 6) ^ ror32(v[4], 11) ^ ror32(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + K[t] + W[t];
			const uint32_t T2 = (ror32(v[0], 2) ^ ror32(v[0], 13) ^ ror32(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
			memmove(v + 1, v, 7 * sizeof(uint32_t));
			v[4] += T1;
			v[0] = T1 + T2;
		}
		for (int i = 0; i < 8; ++i) H[i] += v[i];
	}
	for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)(H[i] >> 24); out[4 * i + 1] = (uint8_t)(H[i] >> 16); out[4 * i + 2] = (uint8_t)(H[i] >> 8); out[4 * i + 3] = (uint8_t)H[i]; }
}
