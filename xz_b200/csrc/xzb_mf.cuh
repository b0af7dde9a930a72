// xzb_mf.cuh -- match finding for one .xz block, restructured for the GPU.
//
// The reference runs the match finder inline with the parser, one position at a time
// (lz/lz_encoder_mf.c).  Its result at position p is a pure function of the block bytes
// [0, p + nice_len) (find and skip do identical insertions), so here it runs as a separate,
// massively parallel pass that fills a per-position "match store" which the parser later
// streams through:
//   1. hash keys for every position (xzb_hash_keys), stable radix sort by (block, key);
//   2. previous occurrence with the same hash-2 / hash-3 / main hash value = the hash heads
//      the reference would have read at p (lz_encoder_mf.c:372-379, 681-688);
//   3. HC: one thread per position walks the chain (hc_find_func :249-287);
//      BT: one thread per hash bucket replays the tree insertions of that bucket's positions
//      in order (bt_find_func :449-512, bt_skip_func :515-568).  Trees of different buckets
//      share no nodes once son[] has one slot per block position instead of being cyclic.
// All positions are block-local; stored positions are p+1 so that 0 is EMPTY_HASH_VALUE.
#pragma once
#include "xzb_common.cuh"

struct __attribute__((aligned(8))) xzb_pair { uint32_t len, dist; };

#define XZB_OVF_MARK 0xFFFFFFFFu

// A stored pair's first word: bits 0-8 match length, bits 9-17 "len2" = how many bytes keep
// matching after skipping one byte behind the match (the reference's len_test_2 for the
// "match + literal + rep0" candidate, lzma_encoder_optimum_normal.c:729-742, computed with the
// limit min(bytes to block end, len + 1 + nice_len); the parser clamps it to its own limit),
// bits 18-25 the byte at match_source[len] (the "match byte" of that literal).
// Both are pure functions of the block bytes, so the match-finder pass can precompute them.
#define XZB_PAIR_LEN(w) ((w) & 0x1FFu)
#define XZB_PAIR_LEN2(w) (((w) >> 9) & 0x1FFu)
#define XZB_PAIR_MB(w) (((w) >> 18) & 0xFFu)

// Per-block view of the match-finder working set (all device pointers).
struct XzbMfBlock {
	const uint8_t *buf;      // block bytes
	uint32_t n;              // block length
	uint32_t room;           // bytes readable from buf (n plus whatever slack follows the block)
	const uint32_t *prev2;   // [n] previous position with same hash-2 value or XZB_NONE (hash_bytes >= 3)
	const uint32_t *prev3;   // [n] same for hash-3 (hash_bytes == 4)
	const uint32_t *prevm;   // [n] same for the main hash (hash chain "son"; HC only)
	uint32_t *son;           // [2n] binary tree children, p+1 encoding (BT only)
	uint32_t *mh;            // [n] match store header: count | longest << 16
	xzb_pair *mp;            // [n * mstride] inline pairs
	xzb_pair *ovf;           // overflow pool of this block
	uint32_t *ovf_top;       // bump pointer (pairs)
	uint32_t ovf_cap;
	uint32_t *err;           // sticky error flag (XZB_MEM_ERROR on pool exhaustion)
};

XZB_HD uint32_t xzb_atomic_add(uint32_t *p, uint32_t v)
{
#ifdef __CUDA_ARCH__
	return atomicAdd(p, v);
#else
	const uint32_t o = *p; *p = o + v; return o;
#endif
}

// hash_2_calc / hash_3_calc / hash_4_calc, lz/lz_encoder_hash.h:53-75.
// Returns the main hash; h2/h3 are the auxiliary hashes (valid when hash_bytes >= 3 / == 4).
XZB_HD uint32_t xzb_hash(const uint8_t *cur, const XzbParams &P, const uint32_t *crc, uint32_t *h2, uint32_t *h3)
{
	if (P.hash_bytes == 2) return (uint32_t)cur[0] | ((uint32_t)cur[1] << 8);
	const uint32_t temp = crc[cur[0]] ^ cur[1];
	*h2 = temp & (XZB_H2_SIZE - 1);
	if (P.hash_bytes == 3) return (temp ^ ((uint32_t)cur[2] << 8)) & P.hash_mask;
	*h3 = (temp ^ ((uint32_t)cur[2] << 8)) & (XZB_H3_SIZE - 1);
	return (temp ^ ((uint32_t)cur[2] << 8) ^ (crc[cur[3]] << 5)) & P.hash_mask;
}

// Writer for one position's pairs: first mstride-1 pairs go straight to the inline slots,
// later ones are parked in `extra` until the final count is known.
struct XzbPairSink {
	xzb_pair *inl;
	uint32_t stride, count;
	xzb_pair extra[XZB_MATCH_LEN_MAX];  // local memory; touched only on overflowing positions
	XZB_HDM void push(uint32_t len, uint32_t dist)
	{
		xzb_pair v; v.len = len; v.dist = dist;
		if (count < stride - 1) inl[count] = v; else extra[count - (stride - 1)] = v;
		++count;
	}
	XZB_HDM void set_len(uint32_t i, uint32_t len)
	{
		if (i < stride - 1) inl[i].len = len; else extra[i - (stride - 1)].len = len;
	}
};

// Finish a position: lzma_mf_find's nice_len extension (lz_encoder_mf.c:45-70), header word,
// spill of pairs beyond the inline slots.
XZB_HD void xzb_mf_finish(const XzbMfBlock &B, const XzbParams &P, uint32_t p, XzbPairSink &S, uint32_t last_len, uint32_t last_dist)
{
	uint32_t longest = 0;
	const uint32_t count = S.count;
	if (count > 0) {
		const uint8_t *p1 = B.buf + p;
		const uint32_t avail1 = B.n - p;  // mf_avail() + 1 after move_pos
		const uint32_t room = B.room - p;
		longest = last_len;
		if (longest == P.nice_len) {
			uint32_t limit = avail1;
			if (limit > XZB_MATCH_LEN_MAX) limit = XZB_MATCH_LEN_MAX;
			longest = xzb_memcmplen_w(p1, p1 - last_dist - 1, longest, limit, room);
		}
		// len2 / match byte of every pair (see XZB_PAIR_* above)
		for (uint32_t i = 0; i < count; ++i) {
			xzb_pair *e = i < S.stride - 1 ? &S.inl[i] : &S.extra[i - (S.stride - 1)];
			const uint32_t L = e->len;
			const uint8_t *bb = p1 - e->dist - 1;
			uint32_t l2 = 0, mb = 0;
			if (L < avail1) {
				mb = bb[L];
				const uint32_t limit = xzb_min(avail1, L + 1 + P.nice_len);
				if (L + 1 < limit) l2 = xzb_memcmplen_w(p1, bb, L + 1, limit, room) - (L + 1);
			}
			e->len = L | (l2 << 9) | (mb << 18);
		}
		if (count >= S.stride) {
			if (count == S.stride) {
				S.inl[S.stride - 1] = S.extra[0];
			} else {
				const uint32_t extra = count - (S.stride - 1);
				const uint32_t at = xzb_atomic_add(B.ovf_top, extra);
				if (at + extra > B.ovf_cap) {
					*B.err = XZB_MEM_ERROR;
				} else {
					for (uint32_t i = 0; i < extra; ++i) B.ovf[at + i] = S.extra[i];
				}
				xzb_pair link; link.len = at; link.dist = XZB_OVF_MARK;
				S.inl[S.stride - 1] = link;
			}
		}
	}
	B.mh[p] = count | (longest << 16);
}

// Head stage shared by all find functions: the hash-2 / hash-3 candidates
// (lzma_mf_hc3_find :303-334, hc4 :365-413, bt3 :619-649, bt4 :674-722).
// Returns len_best for the chain/tree stage; *skip_tree = the reference takes the *_skip()
// path because the head-stage match already reached len_limit.
XZB_HD uint32_t xzb_mf_head(const XzbMfBlock &B, const XzbParams &P, uint32_t p, uint32_t len_limit,
		XzbPairSink &S, uint32_t *last_len, uint32_t *last_dist, bool *skip_tree)
{
	const uint8_t *cur = B.buf + p;
	*skip_tree = false;
	if (P.hash_bytes == 2) return 1;
	const uint32_t q2 = B.prev2[p];
	if (P.hash_bytes == 3) {
		uint32_t len_best = 2;
		if (q2 != XZB_NONE) {
			const uint32_t delta2 = p - q2;
			if (delta2 < P.cyclic_size && *(cur - delta2) == *cur) {
				len_best = xzb_memcmplen_w(cur, cur - delta2, len_best, len_limit, B.room - p);
				S.push(len_best, delta2 - 1);
				*last_len = len_best; *last_dist = delta2 - 1;
				if (len_best == len_limit) *skip_tree = true;
			}
		}
		return len_best;
	}
	const uint32_t q3 = B.prev3[p];
	uint32_t delta2 = q2 != XZB_NONE ? p - q2 : XZB_NONE;
	const uint32_t delta3 = q3 != XZB_NONE ? p - q3 : XZB_NONE;
	uint32_t len_best = 1;
	if (delta2 < P.cyclic_size && *(cur - delta2) == *cur) {
		len_best = 2;
		S.push(2, delta2 - 1);
		*last_len = 2; *last_dist = delta2 - 1;
	}
	if (delta2 != delta3 && delta3 < P.cyclic_size && *(cur - delta3) == *cur) {
		len_best = 3;
		S.push(3, delta3 - 1);
		*last_dist = delta3 - 1;
		delta2 = delta3;
	}
	if (S.count != 0) {
		len_best = xzb_memcmplen_w(cur, cur - delta2, len_best, len_limit, B.room - p);
		S.set_len(S.count - 1, len_best);
		*last_len = len_best;
		if (len_best == len_limit) *skip_tree = true;
	}
	if (len_best < 3) len_best = 3;
	return len_best;
}

// len_limit and the "is this position inserted at all" rule of the header() macro
// (lz_encoder_mf.c:190-201) with the whole block resident (action == LZMA_FINISH at the end).
XZB_HD bool xzb_mf_len_limit(const XzbParams &P, uint32_t n, uint32_t p, uint32_t *len_limit)
{
	const uint32_t avail = n - p;
	if (P.nice_len <= avail) { *len_limit = P.nice_len; return true; }
	if (avail < P.hash_bytes) return false;  // move_pending(): never inserted, zero matches
	*len_limit = avail;
	return true;
}

// ---- hash chain: one thread per position (hc_find_func, lz_encoder_mf.c:249-287) ----
XZB_HD void xzb_hc_position(const XzbMfBlock &B, const XzbParams &P, uint32_t p)
{
	uint32_t len_limit;
	if (!xzb_mf_len_limit(P, B.n, p, &len_limit)) { B.mh[p] = 0; return; }
	XzbPairSink S; S.inl = B.mp + (size_t)p * P.mstride; S.stride = P.mstride; S.count = 0;
	uint32_t last_len = 0, last_dist = 0; bool skip_tree;
	uint32_t len_best = xzb_mf_head(B, P, p, len_limit, S, &last_len, &last_dist, &skip_tree);
	if (!skip_tree) {
		const uint8_t *cur = B.buf + p;
		uint32_t cur_match = B.prevm[p];
		uint32_t depth = P.depth;
		for (;;) {
			if (depth-- == 0 || cur_match == XZB_NONE) break;
			const uint32_t delta = p - cur_match;
			if (delta >= P.cyclic_size) break;
			const uint8_t *pb = cur - delta;
			cur_match = B.prevm[cur_match];
			if (pb[len_best] == cur[len_best] && pb[0] == cur[0]) {
				const uint32_t len = xzb_memcmplen_w(cur, pb, 1, len_limit, B.room - p);
				if (len_best < len) {
					len_best = len;
					S.push(len, delta - 1);
					last_len = len; last_dist = delta - 1;
					if (len == len_limit) break;
				}
			}
		}
	}
	xzb_mf_finish(B, P, p, S, last_len, last_dist);
}

// ---- binary tree: one thread per bucket; positions of the bucket in ascending order ----
// `prev_in_bucket` is the bucket's previous position (the hash head the reference reads), or
// XZB_NONE for the first one.  bt_find_func :449-512 / bt_skip_func :515-568.
XZB_HD void xzb_bt_position(const XzbMfBlock &B, const XzbParams &P, uint32_t p, uint32_t prev_in_bucket)
{
	uint32_t len_limit;
	if (!xzb_mf_len_limit(P, B.n, p, &len_limit)) { B.mh[p] = 0; return; }
	XzbPairSink S; S.inl = B.mp + (size_t)p * P.mstride; S.stride = P.mstride; S.count = 0;
	uint32_t last_len = 0, last_dist = 0; bool skip_tree;
	uint32_t len_best = xzb_mf_head(B, P, p, len_limit, S, &last_len, &last_dist, &skip_tree);
	const uint8_t *cur = B.buf + p;
	uint32_t *son = B.son;
	uint32_t *ptr0 = son + ((size_t)p << 1) + 1;
	uint32_t *ptr1 = son + ((size_t)p << 1);
	uint32_t len0 = 0, len1 = 0;
	uint32_t cur_match = prev_in_bucket == XZB_NONE ? 0 : prev_in_bucket + 1;  // p+1 encoding
	uint32_t depth = P.depth;
	const uint32_t pos = p + 1;
	const uint32_t room = B.room - p;
	for (;;) {
		const uint32_t delta = pos - cur_match;
		if (depth-- == 0 || cur_match == 0 || delta >= P.cyclic_size) {
			*ptr0 = 0; *ptr1 = 0;
			break;
		}
		uint32_t *pair = son + ((size_t)(p - delta) << 1);
		const uint32_t child0 = pair[0], child1 = pair[1];  // issued together with the byte loads below
		const uint8_t *pb = cur - delta;
		uint32_t len = len0 < len1 ? len0 : len1;
		// one word-wise compare starting AT len replaces "pb[len] == cur[len]" + memcmplen(len + 1) and
		// also yields the bytes for the left/right decision below
		uint32_t c_cur = 0, c_pb = 0;
		const uint32_t len_new = xzb_memcmplen_w2(cur, pb, len, len_limit, room, &c_cur, &c_pb);
		if (len_new != len) {
			len = len_new;
			if (!skip_tree) {
				if (len_best < len) {
					len_best = len;
					S.push(len, delta - 1);
					last_len = len; last_dist = delta - 1;
					if (len == len_limit) { *ptr1 = child0; *ptr0 = child1; break; }
				}
			} else if (len == len_limit) {
				*ptr1 = child0; *ptr0 = child1; break;
			}
		}
		if (c_pb < c_cur) {
			*ptr1 = cur_match; ptr1 = pair + 1; cur_match = child1; len1 = len;
		} else {
			*ptr0 = cur_match; ptr0 = pair; cur_match = child0; len0 = len;
		}
	}
	xzb_mf_finish(B, P, p, S, last_len, last_dist);
}
