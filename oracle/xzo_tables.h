/* oracle/xzo_tables.h -- lookup tables of the oracle (TEST INFRASTRUCTURE ONLY).
 * The reference ships these as generated files; we regenerate them at start-up from the
 * published generators: check/crc32_tablegen.c:26-40 (poly 0xEDB88320),
 * check/crc64_tablegen.c:25-40 (poly 0xC96C5795D7870F42),
 * rangecoder/price_tablegen.c:28-56.  lz/lz_encoder_hash.h:30-39: the match-finder
 * hash table is lzma_crc32_table[0]. */
#ifndef XZO_TABLES_H
#define XZO_TABLES_H
#include <stdint.h>
extern uint32_t xzo_crc32_table[256];
extern uint64_t xzo_crc64_table[256];
extern uint8_t xzo_rc_prices[128];
void xzo_tables_init(void);
#endif
