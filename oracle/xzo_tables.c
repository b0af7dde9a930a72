/* oracle/xzo_tables.c -- see xzo_tables.h (TEST INFRASTRUCTURE ONLY). */
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
