// xzb_params.h -- host-side derivation of the per-filter constants (XzbParams) from LZMA2
// options, and the table generators.  Reference: lz/lz_encoder.c:191-368 (lz_encoder_prepare),
// lzma/lzma_encoder.c:453-707 (is_options_valid, set_lz_options, lzma_lzma_encoder_create),
// lzma/lzma_encoder_presets.c:16-63, check/crc32_tablegen.c, check/crc64_tablegen.c,
// rangecoder/price_tablegen.c:28-56.
#pragma once
#include "xzb_common.cuh"
#include "xzb_frame.cuh"

struct XzbLzmaOptions {  // the subset of lzma_options_lzma (api/lzma/lzma12.h:216-525) on the path
	uint32_t dict_size, lc, lp, pb, mode, nice_len, mf, depth;
};

static inline int xzb_preset(XzbLzmaOptions *o, uint32_t preset)  // lzma_lzma_preset
{
	const uint32_t level = preset & 0x1F, flags = preset & ~0x1Fu;
	if (level > 9 || (flags & ~0x80000000u)) return 1;
	static const uint8_t dict_pow2[] = { 18, 20, 21, 22, 22, 23, 23, 24, 25, 26 };
	o->lc = 3; o->lp = 0; o->pb = 2;
	o->dict_size = 1u << dict_pow2[level];
	if (level <= 3) {
		static const uint8_t depths[] = { 4, 8, 24, 48 };
		o->mode = XZB_MODE_FAST;
		o->mf = level == 0 ? XZB_MF_HC3 : XZB_MF_HC4;
		o->nice_len = level <= 1 ? 128 : 273;
		o->depth = depths[level];
	} else {
		o->mode = XZB_MODE_NORMAL;
		o->mf = XZB_MF_BT4;
		o->nice_len = level == 4 ? 16 : level == 5 ? 32 : 64;
		o->depth = 0;
	}
	if (flags & 0x80000000u) {
		o->mode = XZB_MODE_NORMAL;
		o->mf = XZB_MF_BT4;
		if (level == 3 || level == 5) { o->nice_len = 192; o->depth = 0; }
		else { o->nice_len = 273; o->depth = 512; }
	}
	return 0;
}

// Returns XZB_OK or XZB_OPTIONS_ERROR (same validation as the reference's init chain).
static inline int xzb_make_params(const XzbLzmaOptions *o, XzbParams *P)
{
	if (o->lc > 4 || o->lp > 4 || o->lc + o->lp > 4 || o->pb > 4) return XZB_OPTIONS_ERROR;
	if (o->nice_len < XZB_MATCH_LEN_MIN || o->nice_len > XZB_MATCH_LEN_MAX) return XZB_OPTIONS_ERROR;
	if (o->mode != XZB_MODE_FAST && o->mode != XZB_MODE_NORMAL) return XZB_OPTIONS_ERROR;
	if (o->mf != XZB_MF_HC3 && o->mf != XZB_MF_HC4 && o->mf != XZB_MF_BT2 && o->mf != XZB_MF_BT3 && o->mf != XZB_MF_BT4) return XZB_OPTIONS_ERROR;
	// IS_ENC_DICT_SIZE_VALID, lz/lz_encoder.h:22-26: 4 KiB .. 1.5 GiB
	if (o->dict_size < 4096 || o->dict_size > (1u << 30) + (1u << 29)) return XZB_OPTIONS_ERROR;
	P->dict_size = o->dict_size; P->lc = o->lc; P->lp = o->lp; P->pb = o->pb; P->mode = o->mode; P->mf = o->mf;
	P->hash_bytes = o->mf & 0x0F; P->is_bt = (o->mf & 0x10) != 0;
	P->nice_len = o->nice_len > P->hash_bytes ? o->nice_len : P->hash_bytes;
	P->cyclic_size = o->dict_size + 1;
	uint32_t hs;
	if (P->hash_bytes == 2) {
		hs = 0xFFFF;
	} else {
		hs = o->dict_size - 1;
		hs |= hs >> 1; hs |= hs >> 2; hs |= hs >> 4; hs |= hs >> 8;
		hs >>= 1; hs |= 0xFFFF;
		if (hs > (1u << 24)) { if (P->hash_bytes == 3) hs = (1u << 24) - 1; else hs >>= 1; }
	}
	P->hash_mask = hs;
	P->depth = o->depth;
	if (P->depth == 0) P->depth = P->is_bt ? 16 + P->nice_len / 2 : 4 + P->nice_len / 4;
	uint32_t log_size = 0;
	while ((1u << log_size) < o->dict_size) ++log_size;
	P->dist_table_size = log_size * 2;
	P->len_table_size = P->nice_len + 1 - XZB_MATCH_LEN_MIN;
	P->mstride = 8;
	P->dict_prop = xzb_lzma2_dict_prop(o->dict_size);
	P->lclppb = (uint8_t)((o->pb * 5 + o->lp) * 9 + o->lc);
	P->n_pre = 0; P->ff_len = 0;
	return XZB_OK;
}

struct XzbHostTables { uint32_t crc32[256]; uint64_t crc64[256]; uint8_t prices[128]; };

static inline void xzb_make_tables(XzbHostTables *t)
{
	for (uint32_t b = 0; b < 256; ++b) {
		uint32_t r = b;
		for (int i = 0; i < 8; ++i) r = (r & 1) ? (r >> 1) ^ 0xEDB88320u : r >> 1;
		t->crc32[b] = r;
		uint64_t q = b;
		for (int i = 0; i < 8; ++i) q = (q & 1) ? (q >> 1) ^ 0xC96C5795D7870F42ull : q >> 1;
		t->crc64[b] = q;
	}
	for (uint32_t i = 8; i < 2048; i += 16) {
		uint32_t w = i, bit_count = 0;
		for (int j = 0; j < 4; ++j) {
			w *= w; bit_count <<= 1;
			while (w >= (1u << 16)) { w >>= 1; ++bit_count; }
		}
		t->prices[i >> 4] = (uint8_t)((11 << 4) - 15 - bit_count);
	}
}
