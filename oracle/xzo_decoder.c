/*
 * oracle/xzo_decoder.c -- CPU restatement of the reference's .xz/LZMA2 DECODER path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/xzo.h).  One-shot form (whole input and output in
 * memory, as with lzma_code(LZMA_FINISH) on lzma_stream_decoder): the output buffer is the
 * dictionary (no wrap), so dict_repeat/dict_put (lz/lz_decoder.h:181-300) become plain
 * indexing while "full" keeps the reference's meaning for distance validation.
 * Citations are relative to /root/reference/src/liblzma/.
 */
#include "xzo.h"
#include "xzo_tables.h"

#include <stdlib.h>
#include <string.h>

#define STATES 12
#define LIT_STATES 7
#define POS_STATES_MAX 16
#define DIST_SLOTS 64
#define DIST_MODEL_START 4
#define DIST_MODEL_END 14
#define FULL_DISTANCES 128
#define ALIGN_BITS 4

typedef uint16_t prob_t;

/* internal: XZO_OK = finished fine; NEED_INPUT = ran out of input (truncated) */
#define NEED_INPUT 100
#define NEED_OUTPUT 101

typedef struct {
	prob_t choice, choice2, low[POS_STATES_MAX][8], mid[POS_STATES_MAX][8], high[256];
} len_dec_t;

/* lzma_lzma1_decoder, lzma/lzma_decoder.c:106-231 */
typedef struct {
	prob_t literal[16 * 0x300];
	prob_t is_match[STATES][POS_STATES_MAX];
	prob_t is_rep[STATES], is_rep0[STATES], is_rep1[STATES], is_rep2[STATES];
	prob_t is_rep0_long[STATES][POS_STATES_MAX];
	prob_t dist_slot[4][DIST_SLOTS];
	prob_t pos_special[FULL_DISTANCES - DIST_MODEL_END];
	prob_t pos_align[1 << ALIGN_BITS];
	len_dec_t match_len, rep_len;
	uint32_t state, rep0, rep1, rep2, rep3;
	uint32_t pos_mask, lc, literal_mask;
	/* range decoder, rangecoder/range_decoder.h:60-66 */
	uint32_t range, code;
	const uint8_t *in; size_t in_pos, in_end; /* in_end: end of this chunk's compressed bytes */
	int in_truncated; /* the chunk's bytes are cut short by the end of the input */
	int err;
} dec_t;

/* lzma_decoder_reset, lzma/lzma_decoder.c:1034-1114 */
static void dec_reset(dec_t *d, uint32_t lc, uint32_t lp, uint32_t pb)
{
	d->pos_mask = (1u << pb) - 1; d->lc = lc;
	d->literal_mask = (0x100u << lp) - (0x100u >> lc);
	const size_t coders = (size_t)0x300 << (lc + lp);
	for (size_t i = 0; i < coders; ++i) d->literal[i] = 1024;
	d->state = 0; d->rep0 = d->rep1 = d->rep2 = d->rep3 = 0;
	for (int i = 0; i < STATES; ++i) {
		for (uint32_t j = 0; j <= d->pos_mask; ++j) { d->is_match[i][j] = 1024; d->is_rep0_long[i][j] = 1024; }
		d->is_rep[i] = d->is_rep0[i] = d->is_rep1[i] = d->is_rep2[i] = 1024;
	}
	for (int i = 0; i < 4; ++i) for (int j = 0; j < DIST_SLOTS; ++j) d->dist_slot[i][j] = 1024;
	for (int i = 0; i < FULL_DISTANCES - DIST_MODEL_END; ++i) d->pos_special[i] = 1024;
	for (int i = 0; i < (1 << ALIGN_BITS); ++i) d->pos_align[i] = 1024;
	len_dec_t *l[2] = { &d->match_len, &d->rep_len };
	for (int k = 0; k < 2; ++k) {
		l[k]->choice = l[k]->choice2 = 1024;
		for (uint32_t ps = 0; ps < (1u << pb); ++ps) for (int i = 0; i < 8; ++i) { l[k]->low[ps][i] = 1024; l[k]->mid[ps][i] = 1024; }
		for (int i = 0; i < 256; ++i) l[k]->high[i] = 1024;
	}
}

/* rc_normalize, range_decoder.h:144-150; running past the chunk's bytes is an error
 * (lzma2_decoder.c:174-188: in_used > compressed_size -> LZMA_DATA_ERROR), running past
 * the end of a truncated input means "need more input". */
static inline void rc_normalize(dec_t *d)
{
	if (d->range < (1u << 24)) {
		uint8_t b = 0;
		if (d->in_pos < d->in_end) b = d->in[d->in_pos++];
		else if (!d->err) d->err = d->in_truncated ? NEED_INPUT : XZO_DATA_ERROR;
		d->range <<= 8;
		d->code = (d->code << 8) | b;
	}
}

/* rc_if_0 / rc_update_0 / rc_update_1, range_decoder.h:152-214 */
static inline uint32_t rc_bit(dec_t *d, prob_t *prob)
{
	rc_normalize(d);
	const uint32_t bound = (d->range >> 11) * *prob;
	if (d->code < bound) {
		d->range = bound;
		*prob += (2048 - *prob) >> 5;
		return 0;
	}
	d->range -= bound; d->code -= bound;
	*prob -= *prob >> 5;
	return 1;
}

static inline uint32_t rc_bittree(dec_t *d, prob_t *probs, uint32_t bits)
{
	uint32_t s = 1;
	for (uint32_t i = 0; i < bits; ++i) s = (s << 1) | rc_bit(d, &probs[s]);
	return s - (1u << bits);
}

/* len_decode, lzma/lzma_decoder.c:47-97 */
static uint32_t len_decode(dec_t *d, len_dec_t *l, uint32_t pos_state)
{
	if (rc_bit(d, &l->choice) == 0) return 2 + rc_bittree(d, l->low[pos_state], 3);
	if (rc_bit(d, &l->choice2) == 0) return 2 + 8 + rc_bittree(d, l->mid[pos_state], 3);
	return 2 + 16 + rc_bittree(d, l->high, 8);
}

/*
 * One LZMA chunk: lzma_decode, lzma/lzma_decoder.c:234-1021 with uncompressed size known
 * and EOPM not allowed (lzma2_decoder.c:120-123 set_uncompressed(..., false)).
 * out[0..*pos) is the dictionary; dict_full = valid history (lz_decoder.h:147-151);
 * writes exactly `usize` bytes unless an error occurs.
 */
static int lzma_chunk_decode(dec_t *d, uint8_t *out, size_t *pos_ptr, size_t usize,
		size_t dict_start, size_t dict_size_r)
{
	size_t pos = *pos_ptr;
	const size_t limit = pos + usize;
	/* rc_read_init, range_decoder.h:69-91 */
	d->range = UINT32_MAX; d->code = 0;
	for (int i = 0; i < 5; ++i) {
		if (d->in_pos >= d->in_end) { *pos_ptr = pos; return d->in_truncated ? NEED_INPUT : XZO_DATA_ERROR; }
		const uint8_t b = d->in[d->in_pos++];
		if (i == 0 && b != 0x00) { *pos_ptr = pos; return XZO_DATA_ERROR; }
		d->code = (d->code << 8) | b;
	}
	d->err = 0;
	uint32_t state = d->state, rep0 = d->rep0, rep1 = d->rep1, rep2 = d->rep2, rep3 = d->rep3;
	while (pos < limit && !d->err) {
		const uint32_t pos_state = (uint32_t)(pos - dict_start) & d->pos_mask;
		/* NB: dict.pos in the reference counts from the dictionary reset point modulo
		 * alignment (lz_decoder.c:53-63 sets pos = 2*288 = 576, a multiple of 16), so
		 * pos & pos_mask equals (bytes since dict reset) & pos_mask. */
		size_t full = pos - dict_start; if (full > dict_size_r) full = dict_size_r;
		if (rc_bit(d, &d->is_match[state][pos_state]) == 0) {
			const uint32_t prev = pos > dict_start ? out[pos - 1] : 0; /* dict_get0; buf[INIT_POS-1] = 0 */
			prob_t *probs = d->literal + 3u * (((((uint32_t)(pos - dict_start) << 8) + prev) & d->literal_mask) << d->lc);
			uint32_t symbol = 1;
			if (state < LIT_STATES) {
				state = state <= 3 ? 0 : state - 3;
				do { symbol = (symbol << 1) | rc_bit(d, &probs[symbol]); } while (symbol < 0x100);
			} else {
				state = state <= 9 ? state - 3 : state - 6;
				/* rc_matched_literal, range_decoder.h:270-300 */
				uint32_t match_byte = (full > rep0) ? out[pos - rep0 - 1] : 0;
				uint32_t offset = 0x100;
				do {
					match_byte <<= 1;
					const uint32_t match_bit = match_byte & offset;
					const uint32_t bit = rc_bit(d, &probs[offset + match_bit + symbol]);
					symbol = (symbol << 1) | bit;
					if (bit) offset &= match_bit; else offset &= ~match_bit;
				} while (symbol < 0x100);
			}
			out[pos++] = (uint8_t)symbol;
			continue;
		}
		uint32_t len;
		if (rc_bit(d, &d->is_rep[state]) == 0) {
			state = state < LIT_STATES ? 7 : 10;
			rep3 = rep2; rep2 = rep1; rep1 = rep0;
			len = len_decode(d, &d->match_len, pos_state);
			const uint32_t ds = len < 6 ? len - 2 : 3;
			uint32_t slot = rc_bittree(d, d->dist_slot[ds], 6);
			if (slot < DIST_MODEL_START) {
				rep0 = slot;
			} else {
				uint32_t nbits = (slot >> 1) - 1;
				rep0 = 2 | (slot & 1);
				if (slot < DIST_MODEL_END) {
					rep0 <<= nbits;
					prob_t *probs = d->pos_special + rep0 - slot - 1;
					uint32_t sym = 1, off = 1;
					do {
						const uint32_t bit = rc_bit(d, &probs[sym]);
						sym = (sym << 1) | bit;
						if (bit) rep0 += off;
						off <<= 1;
					} while (--nbits > 0);
				} else {
					nbits -= ALIGN_BITS;
					do { /* rc_direct, range_decoder.h:375-388 */
						rc_normalize(d);
						d->range >>= 1;
						d->code -= d->range;
						const uint32_t mask = 0u - (d->code >> 31);
						d->code += d->range & mask;
						rep0 = (rep0 << 1) + (mask + 1);
					} while (--nbits > 0);
					rep0 <<= ALIGN_BITS;
					uint32_t sym = 1, rev = 0;
					for (uint32_t i = 0; i < ALIGN_BITS; ++i) {
						const uint32_t bit = rc_bit(d, &d->pos_align[sym]);
						sym = (sym << 1) | bit; rev |= bit << i;
					}
					rep0 += rev;
					if (rep0 == UINT32_MAX) { d->err = XZO_DATA_ERROR; break; } /* EOPM not allowed in LZMA2 */
				}
			}
			if (!(full > rep0)) { if (!d->err) d->err = XZO_DATA_ERROR; break; }
		} else {
			if (!(full > 0)) { if (!d->err) d->err = XZO_DATA_ERROR; break; }
			if (rc_bit(d, &d->is_rep0[state]) == 0) {
				if (rc_bit(d, &d->is_rep0_long[state][pos_state]) == 0) {
					state = state < LIT_STATES ? 9 : 11;
					/* dict_get(dict, rep0): an invalid rep0 reads stale dictionary bytes in the
					 * reference (lz_decoder.h:181-188); with a freshly reset dictionary rep0 = 0
					 * is the only reachable case and it is valid because full > 0. */
					if (!(full > rep0)) { if (!d->err) d->err = XZO_DATA_ERROR; break; }
					out[pos] = out[pos - rep0 - 1]; ++pos;
					continue;
				}
			} else {
				uint32_t dist;
				if (rc_bit(d, &d->is_rep1[state]) == 0) { dist = rep1; }
				else {
					if (rc_bit(d, &d->is_rep2[state]) == 0) { dist = rep2; }
					else { dist = rep3; rep3 = rep2; }
					rep2 = rep1;
				}
				rep1 = rep0; rep0 = dist;
			}
			state = state < LIT_STATES ? 8 : 11;
			len = len_decode(d, &d->rep_len, pos_state);
			if (!(full > rep0)) { if (!d->err) d->err = XZO_DATA_ERROR; break; }
		}
		if (d->err) break;
		/* dict_repeat, lz_decoder.h:202-266; a match running past the chunk's
		 * uncompressed size is corrupt (lzma_decoder.c:1001-1009). */
		if (len > limit - pos) { d->err = XZO_DATA_ERROR; len = (uint32_t)(limit - pos); }
		const size_t back = pos - rep0 - 1;
		for (uint32_t i = 0; i < len; ++i) out[pos + i] = out[back + i];
		pos += len;
	}
	d->state = state; d->rep0 = rep0; d->rep1 = rep1; d->rep2 = rep2; d->rep3 = rep3;
	*pos_ptr = pos;
	if (d->err) return d->err;
	/* lzma_decoder.c:661-690: one more normalise, then the code must be zero */
	rc_normalize(d);
	if (d->err) return d->err;
	if (d->code != 0) return XZO_DATA_ERROR;
	return XZO_OK;
}

/* lzma2_decode, lzma/lzma2_decoder.c:55-230.
 * out_limit_is_exact: the limit comes from the Block Header's Uncompressed Size, so trying
 * to exceed it is corruption; otherwise it is the caller's capacity (XZO_BUF_ERROR). */
static int lzma2_decode_raw(const uint8_t *in, size_t in_size, uint32_t dict_size,
		uint8_t *out, size_t out_limit, size_t *in_used, size_t *out_used)
{
	dec_t *d = malloc(sizeof(dec_t));
	if (d == NULL) return XZO_MEM_ERROR;
	size_t dict_size_r = dict_size < 4096 ? 4096 : dict_size; /* lz_decoder.c:247-256 */
	dict_size_r = (dict_size_r + 15) & ~(size_t)15;
	size_t in_pos = 0, pos = 0, dict_start = 0;
	int need_properties = 1, need_dictionary_reset = 1;
	uint32_t lc = 0, lp = 0, pb = 0;
	int ret;
	for (;;) {
		if (in_pos >= in_size) { ret = NEED_INPUT; break; }
		const uint32_t control = in[in_pos++];
		if (control == 0x00) { ret = XZO_OK; break; }
		if (control >= 0xE0 || control == 1) {
			need_properties = 1; need_dictionary_reset = 1;
		} else if (need_dictionary_reset) { ret = XZO_DATA_ERROR; break; }
		int is_lzma = control >= 0x80, new_props = 0, state_reset = 0;
		if (is_lzma) {
			if (control >= 0xC0) { need_properties = 0; new_props = 1; }
			else if (need_properties) { ret = XZO_DATA_ERROR; break; }
			else if (control >= 0xA0) state_reset = 1;
		} else if (control > 2) { ret = XZO_DATA_ERROR; break; }
		if (need_dictionary_reset) { need_dictionary_reset = 0; dict_start = pos; }
		size_t usize = 0, csize;
		if (is_lzma) {
			if (in_size - in_pos < 2) { in_pos = in_size; ret = NEED_INPUT; break; }
			usize = ((size_t)(control & 0x1F) << 16) + ((size_t)in[in_pos] << 8) + in[in_pos + 1] + 1;
			in_pos += 2;
		}
		if (in_size - in_pos < 2) { in_pos = in_size; ret = NEED_INPUT; break; }
		csize = ((size_t)in[in_pos] << 8) + in[in_pos + 1] + 1;
		in_pos += 2;
		if (!is_lzma) {
			/* SEQ_COPY: dict_write, lz_decoder.h:283-297 */
			size_t n = csize;
			int short_in = 0, short_out = 0;
			if (n > in_size - in_pos) { n = in_size - in_pos; short_in = 1; }
			if (n > out_limit - pos) { n = out_limit - pos; short_out = 1; short_in = 0; }
			memcpy(out + pos, in + in_pos, n);
			pos += n; in_pos += n;
			if (short_out) { ret = NEED_OUTPUT; break; }
			if (short_in) { ret = NEED_INPUT; break; }
			continue;
		}
		if (new_props) {
			if (in_pos >= in_size) { ret = NEED_INPUT; break; }
			/* lzma_lzma_lclppb_decode, lzma_decoder.c:1198-1211 */
			uint32_t byte = in[in_pos++];
			if (byte > (4 * 5 + 4) * 9 + 8) { ret = XZO_DATA_ERROR; break; }
			pb = byte / (9 * 5); byte -= pb * 9 * 5; lp = byte / 9; lc = byte - lp * 9;
			if (lc + lp > 4) { ret = XZO_DATA_ERROR; break; }
			dec_reset(d, lc, lp, pb);
		} else if (state_reset) {
			dec_reset(d, lc, lp, pb);
		}
		/* SEQ_LZMA :165-196 */
		d->in = in; d->in_pos = in_pos;
		const size_t chunk_start = in_pos;
		const int chunk_cut = csize > in_size - in_pos; /* chunk extends past the bytes we were given */
		d->in_end = chunk_cut ? in_size : in_pos + csize;
		d->in_truncated = chunk_cut;
		size_t want = usize; int short_out = 0;
		if (want > out_limit - pos) { want = out_limit - pos; short_out = 1; }
		ret = lzma_chunk_decode(d, out, &pos, want, dict_start, dict_size_r);
		in_pos = d->in_pos;
		if (short_out && (ret == XZO_OK || ret == XZO_DATA_ERROR)) {
			/* the decoder wanted to write more than the limit allows */
			ret = NEED_OUTPUT; break;
		}
		if (ret != XZO_OK) break;
		/* :190-193: the LZMA decoder finished but compressed bytes are left over */
		if (in_pos - chunk_start != csize) { ret = XZO_DATA_ERROR; break; }
	}
	free(d);
	*in_used = in_pos; *out_used = pos;
	return ret;
}

int xzo_lzma2_decode(const uint8_t *in, size_t in_size, uint32_t dict_size,
		uint8_t *out, size_t out_cap, size_t *out_size, size_t *in_used)
{
	size_t iu = 0, ou = 0;
	int ret = lzma2_decode_raw(in, in_size, dict_size, out, out_cap, &iu, &ou);
	if (out_size) *out_size = ou;
	if (in_used) *in_used = iu;
	if (ret == NEED_INPUT || ret == NEED_OUTPUT) return XZO_BUF_ERROR;
	return ret;
}

/* lzma_vli_decode (single call), common/vli_decoder.c:16-86 */
static int vli_get(const uint8_t *in, size_t *pos, size_t size, uint64_t *v)
{
	*v = 0;
	for (uint32_t i = 0; i < 9; ++i) {
		if (*pos >= size) return NEED_INPUT;
		const uint8_t b = in[(*pos)++];
		*v |= (uint64_t)(b & 0x7F) << (7 * i);
		if ((b & 0x80) == 0) {
			if (b == 0x00 && i != 0) return XZO_DATA_ERROR;
			return XZO_OK;
		}
	}
	return XZO_DATA_ERROR;
}

static uint32_t rd32(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* stream_decode, common/stream_decoder.c:101-378 (flags = 0: one stream, no
 * LZMA_CONCATENATED), driven to completion as lzma_code(LZMA_FINISH) does. */
int xzo_stream_decode(const uint8_t *in, size_t in_size, uint8_t *out, size_t out_cap, size_t *out_size)
{
	xzo_tables_init();
	static const uint8_t magic[6] = { 0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00 };
	size_t ip = 0, op = 0;
	*out_size = 0;
	/* lzma_stream_header_decode, common/stream_flags_decoder.c:26-60 */
	if (in_size < 12) {
		/* magic mismatch is reported as soon as 12 bytes are there; with fewer bytes the
		 * reference waits for more input -> LZMA_BUF_ERROR */
		return XZO_BUF_ERROR;
	}
	if (memcmp(in, magic, 6) != 0) return XZO_FORMAT_ERROR;
	if (xzo_crc32(in + 6, 2, 0) != rd32(in + 8)) return XZO_DATA_ERROR;
	if (in[6] != 0x00 || (in[7] & 0xF0)) return XZO_OPTIONS_ERROR;
	const uint32_t check = in[7] & 0x0F;
	/* lzma_check_size, check/check.c:41-58 */
	static const uint8_t check_sizes[16] = { 0, 4, 4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 32, 64, 64, 64 };
	const uint32_t csize = check_sizes[check];
	ip = 12;
	/* index hash state, common/index_hash.c: we keep the records instead of a SHA-256 */
	size_t nrec = 0, cap = 64;
	uint64_t *recs = malloc(cap * 2 * sizeof(uint64_t));
	if (recs == NULL) return XZO_MEM_ERROR;
	int ret = XZO_OK;
	for (;;) {
		if (ip >= in_size) { ret = XZO_BUF_ERROR; goto done; }
		if (in[ip] == 0x00) break; /* INDEX_INDICATOR */
		/* Block Header: common/block_header_decoder.c:17-124 */
		const uint32_t hsize = ((uint32_t)in[ip] + 1) * 4;
		if (in_size - ip < hsize) { ret = XZO_BUF_ERROR; goto done; }
		const uint8_t *h = in + ip;
		const size_t hin = hsize - 4;
		if (xzo_crc32(h, hin, 0) != rd32(h + hin)) { ret = XZO_DATA_ERROR; goto done; }
		if (h[1] & 0x3C) { ret = XZO_OPTIONS_ERROR; goto done; }
		size_t hp = 2;
		uint64_t comp = UINT64_MAX, uncomp = UINT64_MAX;
		if (h[1] & 0x40) {
			int r = vli_get(h, &hp, hin, &comp);
			if (r != XZO_OK) { ret = r == NEED_INPUT ? XZO_DATA_ERROR : r; goto done; }
			/* lzma_block_unpadded_size() == 0 -> LZMA_DATA_ERROR (block_util.c:53-77) */
			if (comp == 0 || comp > (UINT64_MAX / 2 - 1024 - 64 - 4)) { ret = XZO_DATA_ERROR; goto done; }
		}
		if (h[1] & 0x80) {
			int r = vli_get(h, &hp, hin, &uncomp);
			if (r != XZO_OK) { ret = r == NEED_INPUT ? XZO_DATA_ERROR : r; goto done; }
		}
		const uint32_t nfilters = (h[1] & 3) + 1;
		uint32_t dict_size = 0; int have_lzma2 = 0;
		for (uint32_t f = 0; f < nfilters; ++f) {
			/* lzma_filter_flags_decode, common/filter_flags_decoder.c:16-45 */
			uint64_t id, psize;
			int r = vli_get(h, &hp, hin, &id);
			if (r != XZO_OK) { ret = XZO_DATA_ERROR; goto done; }
			if (id >= (1ull << 62)) { ret = XZO_DATA_ERROR; goto done; }
			r = vli_get(h, &hp, hin, &psize);
			if (r != XZO_OK) { ret = XZO_DATA_ERROR; goto done; }
			if (hin - hp < psize) { ret = XZO_DATA_ERROR; goto done; }
			if (id != 0x21 || nfilters != 1) { ret = XZO_OPTIONS_ERROR; goto done; } /* only the LZMA2-only chain is in scope */
			/* lzma_lzma2_props_decode, lzma/lzma2_decoder.c:298-331 */
			if (psize != 1 || (h[hp] & 0xC0) || h[hp] > 40) { ret = XZO_OPTIONS_ERROR; goto done; }
			dict_size = h[hp] == 40 ? UINT32_MAX : (2u | (h[hp] & 1u)) << (h[hp] / 2u + 11);
			hp += psize; have_lzma2 = 1;
		}
		while (hp < hin) if (h[hp++] != 0x00) { ret = XZO_OPTIONS_ERROR; goto done; }
		if (!have_lzma2) { ret = XZO_OPTIONS_ERROR; goto done; }
		ip += hsize;
		/* Block: common/block_decoder.c:64-200 */
		size_t in_avail = in_size - ip; int truncated = 1;
		if (comp != UINT64_MAX && comp <= in_avail) { in_avail = (size_t)comp; truncated = 0; }
		size_t out_limit = out_cap - op; int out_exact = 0;
		if (uncomp != UINT64_MAX && uncomp <= out_limit) { out_limit = (size_t)uncomp; out_exact = 1; }
		size_t iu = 0, ou = 0;
		int r = lzma2_decode_raw(in + ip, in_avail, dict_size, out + op, out_limit, &iu, &ou);
		if (r == NEED_INPUT) { ret = truncated ? XZO_BUF_ERROR : XZO_DATA_ERROR; goto done; }
		if (r == NEED_OUTPUT) { ret = out_exact ? XZO_DATA_ERROR : XZO_BUF_ERROR; goto done; }
		if (r != XZO_OK) { ret = r; goto done; }
		if ((comp != UINT64_MAX && iu != comp) || (uncomp != UINT64_MAX && ou != uncomp)) { ret = XZO_DATA_ERROR; goto done; }
		const uint8_t *blk_out = out + op;
		ip += iu; op += ou;
		uint64_t c = iu;
		while (c & 3) { /* Block Padding */
			if (ip >= in_size) { ret = XZO_BUF_ERROR; goto done; }
			if (in[ip++] != 0x00) { ret = XZO_DATA_ERROR; goto done; }
			++c;
		}
		if (in_size - ip < csize) { ret = XZO_BUF_ERROR; goto done; }
		if (check == XZO_CHECK_CRC32) { if (xzo_crc32(blk_out, ou, 0) != rd32(in + ip)) { ret = XZO_DATA_ERROR; goto done; } }
		else if (check == XZO_CHECK_CRC64) {
			const uint64_t want = (uint64_t)rd32(in + ip) | ((uint64_t)rd32(in + ip + 4) << 32);
			if (xzo_crc64(blk_out, ou, 0) != want) { ret = XZO_DATA_ERROR; goto done; }
		} else if (check == 10) {
			uint8_t dg[32];
			xzo_sha256(blk_out, ou, dg);
			if (memcmp(dg, in + ip, 32) != 0) { ret = XZO_DATA_ERROR; goto done; }
		}
		ip += csize;
		if (nrec == cap) { cap *= 2; recs = realloc(recs, cap * 2 * sizeof(uint64_t)); if (!recs) return XZO_MEM_ERROR; }
		recs[2 * nrec] = hsize + iu + csize; /* unpadded size */
		recs[2 * nrec + 1] = ou;
		++nrec;
	}
	/* Index: lzma_index_hash_decode, common/index_hash.c:175-341 */
	{
		const size_t istart = ip;
		++ip;
		uint64_t count;
		int r = vli_get(in, &ip, in_size, &count);
		if (r != XZO_OK) { ret = r == NEED_INPUT ? XZO_BUF_ERROR : r; goto done; }
		if (count != nrec) { ret = XZO_DATA_ERROR; goto done; }
		for (size_t i = 0; i < nrec; ++i) {
			uint64_t u, v;
			r = vli_get(in, &ip, in_size, &u);
			if (r != XZO_OK) { ret = r == NEED_INPUT ? XZO_BUF_ERROR : r; goto done; }
			r = vli_get(in, &ip, in_size, &v);
			if (r != XZO_OK) { ret = r == NEED_INPUT ? XZO_BUF_ERROR : r; goto done; }
			if (u != recs[2 * i] || v != recs[2 * i + 1]) { ret = XZO_DATA_ERROR; goto done; }
		}
		while ((ip - istart) & 3) {
			if (ip >= in_size) { ret = XZO_BUF_ERROR; goto done; }
			if (in[ip++] != 0x00) { ret = XZO_DATA_ERROR; goto done; }
		}
		if (in_size - ip < 4) { ret = XZO_BUF_ERROR; goto done; }
		if (xzo_crc32(in + istart, ip - istart, 0) != rd32(in + ip)) { ret = XZO_DATA_ERROR; goto done; }
		ip += 4;
		const size_t isize = ip - istart;
		/* Stream Footer: common/stream_flags_decoder.c:63-88, stream_decoder.c:296-332 */
		if (in_size - ip < 12) { ret = XZO_BUF_ERROR; goto done; }
		const uint8_t *f = in + ip;
		if (f[10] != 'Y' || f[11] != 'Z') { ret = XZO_DATA_ERROR; goto done; } /* FORMAT_ERROR -> DATA_ERROR */
		if (xzo_crc32(f + 4, 6, 0) != rd32(f)) { ret = XZO_DATA_ERROR; goto done; }
		if (f[8] != 0x00 || (f[9] & 0xF0)) { ret = XZO_OPTIONS_ERROR; goto done; }
		if (((uint64_t)rd32(f + 4) + 1) * 4 != isize) { ret = XZO_DATA_ERROR; goto done; }
		if ((f[9] & 0x0F) != check) { ret = XZO_DATA_ERROR; goto done; }
		ip += 12;
		/* flags = 0 (no LZMA_CONCATENATED): LZMA_STREAM_END right after the first Stream's
		 * footer; any bytes after it are left unread (stream_decoder.c:334-336). */
	}
done:
	free(recs);
	*out_size = op;
	return ret;
}
