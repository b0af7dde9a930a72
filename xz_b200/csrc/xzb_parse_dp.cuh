// xzb_parse_dp.cuh -- normal-mode parser (lzma_lzma_optimum_normal) as a forward DP whose
// serial chain is a handful of register operations per position (device only).
//
// Same decisions, bit for bit, as lzma_encoder_optimum_normal.c:270-858.  What is different from
// both the reference and the three-warp kernel in xzb_parse_warp.cuh is the organisation:
//
//   * Candidates are "pushed" into a 1024-entry ring of 16-byte slots (price, back, packed link,
//     back_2) indexed by DP node; a node's slot is final once every earlier node has pushed.
//     A slot also carries the byte buf[target - rep0 - 1] ("match byte" of the target node), so
//     finishing a node needs no window access before its literal can be priced.
//   * Everything that depends only on the block position -- match list, get_dist_len_price() of
//     every length, the whole "match + literal + rep0" candidate except two state bits, the plain
//     literal price, the match bytes of all match targets -- comes from the helper warp, which runs
//     up to 16 positions ahead inside the segment (prices are frozen while a segment's DP runs).
//   * Everything that depends on the node's state arrives through one 32-byte "price bundle"
//     pb[state][pos_state] rebuilt once per segment: is_match / is_rep / is_rep0 / is_rep0_long /
//     is_rep1 / is_rep2 bit prices pre-added in the seven combinations helper2 uses.
//   * Per node the DP warp then does: link -> state/reps (one 16-byte node record), one round of
//     window loads for the four reps (lane = rep x byte), the literal / short-rep step into
//     slot cur+1, and the pushes: one lane per candidate length, classes applied in the
//     reference's program order so that equal prices resolve identically (strict '<' keeps the
//     first).
//   * backward() turns the links into a symbol stack; encode_symbol / the range coder / the LZMA2
//     chunker are the warp-cooperative ones of WarpEncT (xzb_parse_warp.cuh).
//
// lzma_encoder_optimum_normal.c line references are given per function.
#pragma once
#include "xzb_parse_warp.cuh"

#define DP_RING 1024u                 // > 2 * XZB_MATCH_LEN_MAX + 1 (furthest candidate of a node)
#define DP_RMASK (DP_RING - 1u)
#define DP_HR_MAX 16u                 // helper record ring (positions ahead of the DP warp)
#define DP_MAXM 32u                   // matches per record; further ones are priced by the DP warp itself
#define DP_PLAIN_POOL 2176u           // uint2 entries shared by the records' per-length tables
#define DP_STALL_HDR 0xFFFFFFFFu

// packed link of a slot / final node: d1 = target - pos_prev (1..273), flags bit0 prev_1_is_literal,
// bit1 prev_2, d2 = pos_prev - pos_prev_2 (x + literal: len_x + 1), mb = byte at target - rep0 - 1
#define DP_META(d1, flags, d2, mb) ((d1) | ((flags) << 9) | ((d2) << 11) | ((mb) << 20))
#define DP_D1(m) ((m) & 0x1FFu)
#define DP_FLAGS(m) (((m) >> 9) & 3u)
#define DP_D2(m) (((m) >> 11) & 0x1FFu)
#define DP_MB(m) (((m) >> 20) & 0xFFu)

struct DpRec {                        // state-independent facts of one block position (helper warp -> DP warp)
	volatile uint32_t tag;            // ((epoch << 16) | node) + 1 once complete
	uint32_t hdr;                     // count | longest << 16 (match store header), DP_STALL_HDR = watchdog
	uint32_t bytes;                   // buf[p] | buf[p-1] << 8
	uint32_t lit_plain;               // get_literal_price(..., match_mode = false, ...)
	uint32_t m_pack[DP_MAXM];         // len | len_test_2 << 9 | (byte at target - dist - 1) << 18
	uint32_t m_dist[DP_MAXM];
	uint32_t m_rel[DP_MAXM];          // "match + literal + rep0" price minus (normal_match_price + is_match[state_after_match] bit 0)
};

struct DS {  // dynamic shared memory of xzb_k_parse_dp
	// ---- coder state shared with WarpEncT's methods ----
	uint32_t len_prices[2][XZB_POS_STATES_MAX][XZB_LEN_SYMBOLS];
	uint32_t dist_slot_prices[XZB_DIST_STATES][XZB_DIST_SLOTS];
	uint32_t dist_prices[XZB_DIST_STATES][XZB_FULL_DISTANCES];
	uint32_t align_prices[XZB_ALIGN_SIZE];
	uint32_t len_counters[2][XZB_POS_STATES_MAX];
	alignas(16) xzb_pair ring_mp[32][8];
	uint32_t ring_mh[32];
	uint32_t m_dist[XZB_MATCH_LEN_MAX + 1];
	uint16_t m_len[XZB_MATCH_LEN_MAX + 1], m_len2[XZB_MATCH_LEN_MAX + 1];
	uint8_t m_mb[XZB_MATCH_LEN_MAX + 1 + 2];
	xzb_prob probs[PI_TOTAL + 2];
	uint8_t prices[128];
	// ---- DP ----
	alignas(16) uint4 slot[DP_RING];     // candidate ring: x price, y back_prev, z DP_META, w back_prev_2
	uint4 n_reps[DP_RING];               // reps[] of finished nodes (ring); slot+n_reps double as the symbol stack
	uint8_t n_st[DP_RING];               // state of finished nodes (ring)
	uint32_t o_back[XZB_OPTS], o_meta[XZB_OPTS], o_back2[XZB_OPTS];   // final links, read by backward()
	alignas(16) uint4 pb[XZB_STATES][XZB_POS_STATES_MAX][2];  // price bundles, see build_bundles()
	// ---- helper warp ----
	alignas(16) DpRec rec[DP_HR_MAX];
	alignas(8) uint2 plain_pool[DP_PLAIN_POOL];   // per record: [len - 2] = { get_dist_len_price | mb << 16, dist }
	alignas(16) xzb_pair mring_mp[32][8];
	uint32_t mring_mh[32];
	volatile uint32_t h_epoch, h_pos0, h_position0, h_consumed, m_exit;
};

struct DpEnc : WarpEncT<DS> {
	uint32_t hr_mask, plain_stride;   // helper ring geometry (depends on nice_len)
	uint32_t sym_cur, sym_end;        // symbol stack [sym_cur, sym_end) left over from the last backward()
	uint32_t *trace; uint32_t trace_cap, trace_n;

	__device__ DpEnc(DS &s, uint32_t l) : WarpEncT<DS>(s, l) {}

	__device__ void reset() { WarpEncT<DS>::reset(); sym_cur = sym_end = 0; }  // lzma_lzma_encoder_reset: opts_*_index = 0
	__device__ __forceinline__ void fast_restart(uint32_t) {}   // fast mode never runs on this kernel
	__device__ __forceinline__ uint2 *sym_stack() const { return reinterpret_cast<uint2 *>(&S.slot[0]); }  // 4096 x {back, len}
	__device__ __forceinline__ uint2 *plain_of(uint32_t node) const { return &S.plain_pool[(node & hr_mask) * plain_stride]; }
	static __device__ __forceinline__ uint32_t st_lit(uint32_t s) { return s <= 3 ? 0 : (s <= 9 ? s - 3 : s - 6); }

	// Bit prices of the state-dependent flags, pre-added the way helper1/helper2 use them (:329-340, :520-548, :602-618):
	// [0] = { is_match 0, is_match 1 + is_rep 0 (normal match), is_match 1 + is_rep 1 (rep match), + short rep }
	// [1] = rep match + get_pure_rep_price(0..3)
	__device__ void build_bundles()
	{
		for (uint32_t t = lane; t < XZB_STATES * num_pos_states; t += 32) {
			const uint32_t st = t / num_pos_states, ps = t - st * num_pos_states;
			const uint32_t m0 = pr0(PI_IS_MATCH + (st << 4) + ps), m1 = pr1(PI_IS_MATCH + (st << 4) + ps);
			const uint32_t e0 = pr0(PI_IS_REP + st), e1 = pr1(PI_IS_REP + st);
			const uint32_t g0 = pr0(PI_IS_REP0 + st), g1 = pr1(PI_IS_REP0 + st);
			const uint32_t l0 = pr0(PI_IS_REP0_LONG + (st << 4) + ps), l1 = pr1(PI_IS_REP0_LONG + (st << 4) + ps);
			const uint32_t h0 = pr0(PI_IS_REP1 + st), h1 = pr1(PI_IS_REP1 + st);
			const uint32_t k0 = pr0(PI_IS_REP2 + st), k1 = pr1(PI_IS_REP2 + st);
			const uint32_t R = m1 + e1;
			S.pb[st][ps][0] = make_uint4(m0, m1 + e0, R, R + g0 + l0);
			S.pb[st][ps][1] = make_uint4(R + g0 + l1, R + g1 + h0, R + g1 + h1 + k0, R + g1 + h1 + k1);
		}
		__syncwarp();
	}
	__device__ __forceinline__ uint32_t bundle_rep(const uint4 &b1, uint32_t r) const { return r == 0 ? b1.x : r == 1 ? b1.y : r == 2 ? b1.z : b1.w; }

	// one lane per candidate; targets of the valid lanes are distinct (strict '<': an earlier candidate keeps the slot)
	__device__ __forceinline__ void push(bool valid, uint32_t target, uint32_t price, uint32_t back, uint32_t meta, uint32_t back2)
	{
		if (valid) {
			uint4 *s = &S.slot[target & DP_RMASK];
			if (price < s->x) *s = make_uint4(price, back, meta, back2);
		}
	}

	// ---- backward (:222-263): links -> symbol stack, first symbol returned ----
	__device__ void backward(uint32_t *len_res, uint32_t *back_res, uint32_t end)
	{
		__syncwarp();
		uint32_t k = XZB_OPTS;
		if (lane == 0) {
			// the walk only reads o_*; the stack overlays slot/n_reps, which the DP no longer needs
			uint2 *stk = sym_stack();
			uint32_t c = end;
			while (c != 0) {
				const uint32_t meta = S.o_meta[c], back = S.o_back[c];
				const uint32_t d1 = DP_D1(meta), fl = DP_FLAGS(meta);
				stk[--k] = make_uint2(back, d1);
				uint32_t pp = c - d1;
				if (fl & 1) {
					stk[--k] = make_uint2(XZB_BACK_LITERAL, 1);
					pp -= 1;
					if (fl & 2) {
						const uint32_t xl = DP_D2(meta) - 1;
						stk[--k] = make_uint2(S.o_back2[c], xl);
						pp -= xl;
					}
				}
				c = pp;
			}
		}
		k = __shfl_sync(WFULL, k, 0);
		__syncwarp();
		const uint2 first = sym_stack()[k];
		*back_res = first.x; *len_res = first.y;
		sym_cur = k + 1; sym_end = XZB_OPTS;
	}

	__device__ void ring_clear()
	{
		for (uint32_t i = lane; i < DP_RING; i += 32) S.slot[i] = make_uint4(XZB_INFINITY_PRICE, 0, 0, 0);
		__syncwarp();
	}

	// ---- helper1 (:270-439): node 0 of a segment.  Returns len_end or 0xFFFFFFFF when the symbol is decided. ----
	__device__ uint32_t helper1(uint32_t *back_res, uint32_t *len_res, uint32_t position)
	{
		uint32_t len_main, mcount;
		if (read_ahead == 0) {
			len_main = mf_find(&mcount);
		} else {
			len_main = longest_match_length;
			mcount = matches_count;
		}
		if (mf_stalled) { *back_res = XZB_BACK_LITERAL; *len_res = 1; return 0xFFFFFFFFu; }
		const uint32_t buf_avail = xzb_min(mf_avail() + 1, XZB_MATCH_LEN_MAX);
		if (buf_avail < 2) { *back_res = XZB_BACK_LITERAL; *len_res = 1; return 0xFFFFFFFFu; }
		const uint32_t p0 = read_pos - 1;
		const uint8_t *b = buf + p0;
		uint32_t rl[4];
		rep_lens4(b, buf_avail, rl);
		uint32_t rep_max_index = 0;
		for (uint32_t i = 1; i < XZB_REPS; ++i) if (rl[i] > rl[rep_max_index]) rep_max_index = i;
		if (rl[rep_max_index] >= nice_len) {
			*back_res = rep_max_index; *len_res = rl[rep_max_index];
			mf_skip(*len_res - 1); return 0xFFFFFFFFu;
		}
		if (len_main >= nice_len) {
			*back_res = S.m_dist[mcount - 1] + XZB_REPS; *len_res = len_main;
			mf_skip(len_main - 1); return 0xFFFFFFFFu;
		}
		const uint32_t current_byte = b[0];
		const uint32_t match_byte = *(b - rep0 - 1);
		if (len_main < 2 && current_byte != match_byte && rl[rep_max_index] < 2) {
			*back_res = XZB_BACK_LITERAL; *len_res = 1; return 0xFFFFFFFFu;
		}
		build_bundles();   // probabilities are frozen from here to the end of the segment
		const uint32_t pos_state = position & pos_mask;
		const uint4 b0 = S.pb[state][pos_state][0], b1 = S.pb[state][pos_state][1];
		const uint32_t lit = literal_price(position, b[-1], state >= XZB_LIT_STATES, match_byte, current_byte);
		uint32_t price1 = b0.x + lit;
		uint32_t back1 = XZB_BACK_LITERAL;
		if (match_byte == current_byte) {
			const uint32_t srp = b0.w;
			if (srp < price1) { price1 = srp; back1 = 0; }
		}
		const uint32_t len_end = xzb_max(len_main, rl[rep_max_index]);
		if (len_end < 2) { *back_res = back1; *len_res = 1; return 0xFFFFFFFFu; }
		ring_clear();
		if (lane == 0) {
			S.n_st[0] = (uint8_t)state;
			S.n_reps[0] = make_uint4(rep0, rep1, rep2, rep3);
			S.slot[1] = make_uint4(price1, back1, DP_META(1u, 0u, 0u, (uint32_t)*(b + 1 - rep0 - 1)), 0);
		}
		__syncwarp();
		for (uint32_t i = 0; i < XZB_REPS; ++i) {
			const uint32_t rep_len = rl[i];
			if (rep_len < 2) continue;
			const uint32_t price = bundle_rep(b1, i);
			const uint32_t rr = rep_of(i);
			for (uint32_t l = 2 + lane; l - lane <= rep_len; l += 32) {
				const bool v = l <= rep_len;
				uint32_t p = 0, mbt = 0;
				if (v) { p = price + len_price(1, l, pos_state); mbt = *(b + l - rr - 1); }
				push(v, l, p, i, DP_META(l, 0u, 0u, mbt), 0);
			}
			__syncwarp();
		}
		const uint32_t start = rl[0] >= 2 ? rl[0] + 1 : 2;
		if (start <= len_main) {
			for (uint32_t l = start + lane; l - lane <= len_main; l += 32) {
				const bool v = l <= len_main;
				uint32_t p = 0, dist = 0, mbt = 0;
				if (v) {
					const uint32_t i = match_index_for(l, mcount);
					dist = S.m_dist[i];
					p = b0.y + dist_len_price(dist, l, pos_state);
					mbt = *(b + l - dist - 1);
				}
				push(v, l, p, dist + XZB_REPS, DP_META(l, 0u, 0u, mbt), 0);
			}
			__syncwarp();
		}
		return len_end;
	}

	// One "X + literal + rep0" candidate evaluated by the whole warp (reps: :635-687).  price_x = price up to
	// and including X (a rep of length len_test from node cur), st_x = state after X.
	__device__ void xlr_push(uint32_t price_x, uint32_t st_x, const uint8_t *b, const uint8_t *bb, uint32_t len_test, uint32_t lt2,
			uint32_t position, uint32_t cur, uint32_t back_x, uint32_t &len_end)
	{
		uint32_t psn = (position + len_test) & pos_mask;
		const uint32_t calp = price_x + S.pb[st_x][psn][0].x
				+ literal_price(position + len_test, b[len_test - 1], true, bb[len_test], b[len_test]);
		const uint32_t st2 = st_lit(st_x);
		psn = (position + len_test + 1) & pos_mask;
		const uint32_t p = calp + S.pb[st2][psn][1].x + len_price(1, lt2, psn);
		const uint32_t offset = cur + len_test + 1 + lt2;
		len_end = xzb_max(len_end, offset);
		const uint32_t mbt = bb[len_test + 1 + lt2];
		__syncwarp();
		push(lane == 0, offset, p, 0, DP_META(lt2, 3u, len_test + 1, mbt), back_x);
		__syncwarp();
	}

	// ---- lzma_lzma_optimum_normal (:802-858) with helper2 (:442-799) inlined in push form ----
	__device__ void optimum_normal(uint32_t *back_res, uint32_t *len_res, uint32_t position)
	{
		if (sym_cur != sym_end) {
			const uint2 s = sym_stack()[sym_cur++];
			*back_res = s.x; *len_res = s.y;
			return;
		}
		if (read_ahead == 0) {
			if (match_price_count >= (1 << 7)) fill_dist_prices();
			if (align_price_count >= XZB_ALIGN_SIZE) fill_align_prices();
		}
		uint32_t len_end = helper1(back_res, len_res, position);
		if (len_end == 0xFFFFFFFFu) return;
		// node c sits at block position P0 + c, LZMA position position + c
		const uint32_t P0 = read_pos - 1;
		const uint32_t epoch = (S.h_epoch + 1) & 0x7FFF;
		__syncwarp();
		if (lane < DP_HR_MAX) S.rec[lane].tag = 0;
		__syncwarp();
		if (lane == 0) {
			S.h_pos0 = P0; S.h_position0 = position; S.h_consumed = 0;
			__threadfence_block();
			S.h_epoch = epoch;
		}
		__syncwarp();

		uint32_t cur;
		for (cur = 1; cur < len_end; ++cur) {
			// ---- the helper's record of this position (mf_find equivalent) ----
			const DpRec *R = &S.rec[cur & hr_mask];
			{
				const uint32_t want = ((epoch << 16) | cur) + 1;
				while (R->tag != want) { }
				__threadfence_block();
			}
			const uint32_t hdr = R->hdr;
			if (hdr == DP_STALL_HDR) { mf_stalled = true; break; }
			const uint32_t mcount = hdr & 0xFFFF, longest = hdr >> 16;
			matches_count = mcount; longest_match_length = longest;
			++read_pos; ++read_ahead;
			if (longest >= nice_len) {  // :846-847; the next call's helper1 wants the match list itself
				--read_pos; --read_ahead;
				mf_find(&matches_count);
				break;
			}
			const uint32_t p = P0 + cur;
			const uint32_t pos = position + cur;
			const uint32_t ps = pos & pos_mask;
			const uint32_t baf = xzb_min(size - p, XZB_OPTS - 1 - cur);   // buf_avail_full
			const uint8_t *b = buf + p;
			const uint32_t cb = R->bytes & 0xFF;

			// ---- node cur: link -> state, reps (:453-497) ----
			const uint4 W = S.slot[cur & DP_RMASK];
			const uint32_t meta = W.z;
			const uint32_t d1 = DP_D1(meta), fl = DP_FLAGS(meta);
			const uint32_t src = fl == 0 ? cur - d1 : (fl == 1 ? cur - d1 - 1 : cur - d1 - DP_D2(meta));
			const uint32_t st_src = S.n_st[src & DP_RMASK];
			const uint4 rs = S.n_reps[src & DP_RMASK];
			uint32_t st, r0, r1, r2, r3;
			{
				const uint32_t xb = fl == 3 ? W.w : W.y;   // the symbol that last changed the reps
				if (fl == 0 && d1 == 1) {                  // literal or short rep from cur - 1
					st = xb == 0 ? (st_src < XZB_LIT_STATES ? 9u : 11u) : st_lit(st_src);
					r0 = rs.x; r1 = rs.y; r2 = rs.z; r3 = rs.w;
				} else if (fl == 1) {                      // literal + rep0: reps as at the source
					st = 8u;
					r0 = rs.x; r1 = rs.y; r2 = rs.z; r3 = rs.w;
				} else {
					st = fl == 3 ? 8u : (xb < XZB_REPS ? (st_src < XZB_LIT_STATES ? 8u : 11u) : (st_src < XZB_LIT_STATES ? 7u : 10u));
					if (xb < XZB_REPS) {
						if (xb == 0) { r0 = rs.x; r1 = rs.y; r2 = rs.z; r3 = rs.w; }
						else if (xb == 1) { r0 = rs.y; r1 = rs.x; r2 = rs.z; r3 = rs.w; }
						else if (xb == 2) { r0 = rs.z; r1 = rs.x; r2 = rs.y; r3 = rs.w; }
						else { r0 = rs.w; r1 = rs.x; r2 = rs.y; r3 = rs.z; }
					} else {
						r0 = xb - XZB_REPS; r1 = rs.x; r2 = rs.y; r3 = rs.z;
					}
				}
			}
			const uint32_t mb = DP_MB(meta);           // = buf[p - r0 - 1]
			const uint32_t cur_price = W.x;
			__syncwarp();
			if (lane == 0) {
				S.n_st[cur & DP_RMASK] = (uint8_t)st;
				S.n_reps[cur & DP_RMASK] = make_uint4(r0, r1, r2, r3);
				S.o_back[cur] = W.y; S.o_meta[cur] = meta; S.o_back2[cur] = W.w;
				S.slot[cur & DP_RMASK].x = XZB_INFINITY_PRICE;   // ring entry is free for node cur + DP_RING
			}
			const uint4 b0 = S.pb[st][ps][0], b1 = S.pb[st][ps][1];

			// ---- one round of window loads for the rep phase: lane = (rep index, byte 0..7) ----
			const uint32_t buf_avail = xzb_min(baf, nice_len);
			const uint32_t hr[4] = { r0, r1, r2, r3 };
			uint32_t rmask;
			{
				const uint32_t j = lane & 7;
				const uint32_t rr = hr[lane >> 3];
				const bool in = j < buf_avail;
				const uint32_t av = in ? b[j] : 0u, cv = in ? (b - rr - 1)[j] : 0x100u;
				rmask = __ballot_sync(WFULL, av != cv);
			}

			// ---- literal and short rep into slot cur + 1 (:499-548) ----
			const uint32_t lit = st < XZB_LIT_STATES ? R->lit_plain : literal_price(pos, R->bytes >> 8, true, mb, cb);
			const uint32_t c1 = cur_price + b0.x + lit;      // cur_and_1_price
			bool next_is_literal = false;
			{
				uint4 N = S.slot[(cur + 1) & DP_RMASK];
				bool dirty = false;
				const uint32_t mb1 = baf >= 2 ? (uint32_t)*(b - r0) : 0u;   // buf[(p + 1) - r0 - 1]
				if (c1 < N.x) { N = make_uint4(c1, XZB_BACK_LITERAL, DP_META(1u, 0u, 0u, mb1), 0); dirty = true; next_is_literal = true; }
				if (mb == cb && !(DP_D1(N.z) > 1 && N.y == 0)) {
					const uint32_t srp = cur_price + b0.w;
					if (srp <= N.x) { N = make_uint4(srp, 0, DP_META(1u, 0u, 0u, mb1), 0); dirty = true; next_is_literal = true; }
				}
				if (dirty) { __syncwarp(); if (lane == 0) S.slot[(cur + 1) & DP_RMASK] = N; }
				__syncwarp();
			}
			if (baf < 2) { if (lane == 0) S.h_consumed = cur; continue; }

			// ---- literal + rep0 (:562-597) ----
			if (!next_is_literal && mb != cb) {
				const uint8_t *bb = b - r0 - 1;
				const uint32_t limit = xzb_min(baf, nice_len + 1);
				const uint32_t len_test = mlen_from(rmask & 0xFF, 1, buf_avail, b, bb, limit) - 1;
				if (len_test >= 2) {
					const uint32_t st2 = st_lit(st);
					const uint32_t psn = (pos + 1) & pos_mask;
					const uint32_t pr_ = c1 + S.pb[st2][psn][1].x + len_price(1, len_test, psn);
					const uint32_t offset = cur + 1 + len_test;
					len_end = xzb_max(len_end, offset);
					push(lane == 0, offset, pr_, 0, DP_META(len_test, 1u, 0u, (uint32_t)bb[1 + len_test]), 0);
					__syncwarp();
				}
			}

			// ---- rep matches (:602-688) ----
			uint32_t start_len = 2;
#pragma unroll
			for (uint32_t ri = 0; ri < XZB_REPS; ++ri) {
				const uint32_t mg = (rmask >> (8 * ri)) & 0xFF;
				if (mg & 3) continue;   // not_equal_16
				const uint8_t *bb = b - hr[ri] - 1;
				const uint32_t len_test = mlen_from(mg, 2, buf_avail, b, bb, buf_avail);
				len_end = xzb_max(len_end, cur + len_test);
				const uint32_t price = cur_price + bundle_rep(b1, ri);
				for (uint32_t l = 2 + lane; l - lane <= len_test; l += 32) {
					const bool v = l <= len_test;
					uint32_t pp = 0, mbt = 0;
					if (v) { pp = price + len_price(1, l, ps); mbt = bb[l]; }
					push(v, cur + l, pp, ri, DP_META(l, 0u, 0u, mbt), 0);
				}
				__syncwarp();
				if (ri == 0) start_len = len_test + 1;
				uint32_t lt2 = len_test + 1;
				const uint32_t limit = xzb_min(baf, lt2 + nice_len);
				if (lt2 < limit) lt2 = mlen_from(mg, lt2, buf_avail, b, bb, limit);
				lt2 -= len_test + 1;
				if (lt2 >= 2)
					xlr_push(price + len_price(1, len_test, ps), st < XZB_LIT_STATES ? 8u : 11u, b, bb, len_test, lt2, pos, cur, ri, len_end);
			}

			// ---- normal matches (:690-796) ----
			const uint32_t new_len = xzb_min(longest, buf_avail);   // :692-700 (the shortened last match has no X+literal+rep0 candidate)
			if (new_len >= start_len) {
				const uint32_t nmp = cur_price + b0.y;               // normal_match_price
				const uint32_t s2 = st < XZB_LIT_STATES ? 7u : 10u;
				len_end = xzb_max(len_end, cur + new_len);
				// "match + literal + rep0" of every match, in match order (they come before the plain
				// candidate of the same slot)
				for (uint32_t base = 0; base < mcount; base += 32) {
					const uint32_t i = base + lane;
					uint32_t off = 0, pp = 0, L = 0, dist = 0, lt2 = 0, mbt = 0;
					bool valid = false;
					if (i < mcount) {
						uint32_t pk, rel;
						if (i < DP_MAXM) { pk = R->m_pack[i]; dist = R->m_dist[i]; rel = R->m_rel[i]; }
						else mlr_eval(p, pos, ps, match_pair(p, i), pk, dist, rel);
						L = pk & 0x1FF; lt2 = (pk >> 9) & 0x1FF; mbt = pk >> 18;
						if (L >= start_len && lt2 >= 2) {
							if (baf < L + 1 + lt2) {   // the DP window (or the block) ends inside the rep0 part: shorter rep0
								const uint32_t n2 = baf > L + 1 ? baf - (L + 1) : 0;
								if (n2 >= 2) {
									const uint32_t psn = (pos + L + 1) & pos_mask;
									rel = rel - len_price(1, lt2, psn) + len_price(1, n2, psn);
									mbt = *(b + L + 1 + n2 - dist - 1);
								}
								lt2 = n2;
							}
							if (lt2 >= 2) {
								valid = true;
								pp = nmp + rel + S.pb[s2][(pos + L) & pos_mask][0].x;
								off = cur + L + 1 + lt2;
							}
						}
					}
					const uint32_t vm = __ballot_sync(WFULL, valid);
					if (vm) {
						const uint32_t same = __match_any_sync(WFULL, valid ? off : 0xFFFFFFFFu - lane);
						const bool clash = __any_sync(WFULL, valid && (same & (same - 1)) != 0);
						len_end = xzb_max(len_end, __reduce_max_sync(WFULL, valid ? off : 0u));
						if (!clash) {
							push(valid, off, pp, 0, DP_META(lt2, 3u, L + 1, mbt), dist + XZB_REPS);
							__syncwarp();
						} else {
							uint32_t todo = vm;
							while (todo) {   // two candidates want the same slot: one at a time, in match order
								const uint32_t j = (uint32_t)__ffs((int)todo) - 1;
								todo &= todo - 1;
								push(lane == j, off, pp, 0, DP_META(lt2, 3u, L + 1, mbt), dist + XZB_REPS);
								__syncwarp();
							}
						}
					}
				}
				// plain matches, one lane per length
				const uint2 *PL = plain_of(cur);
				for (uint32_t l = start_len + lane; l - lane <= new_len; l += 32) {
					const bool v = l <= new_len;
					uint32_t pp = 0, dist = 0, mbt = 0;
					if (v) { const uint2 e = PL[l - 2]; pp = nmp + (e.x & 0xFFFF); mbt = e.x >> 16; dist = e.y; }
					push(v, cur + l, pp, dist + XZB_REPS, DP_META(l, 0u, 0u, mbt), 0);
				}
				__syncwarp();
			}
			if (lane == 0) S.h_consumed = cur;
		}
		// the end node's link (its slot is final: every earlier node has pushed)
		{
			const uint4 W = S.slot[cur & DP_RMASK];
			__syncwarp();
			if (lane == 0) { S.o_back[cur] = W.y; S.o_meta[cur] = W.z; S.o_back2[cur] = W.w; }
		}
		backward(len_res, back_res, cur);
	}

	// the i-th (len, dist) pair of block position p straight from the match store (i >= DP_MAXM only)
	__device__ __forceinline__ xzb_pair match_pair(uint32_t p, uint32_t i) const
	{
		const xzb_pair *inl = g_mp + (size_t)p * 8;
		const uint32_t count = g_mh[p] & 0xFFFF;
		if (count <= 8 || i < 7) return inl[i];
		const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(g_ovf + inl[7].len + (i - 7)));
		return xzb_pair{ v.x, v.y };
	}

	// State-independent part of the "match + literal + rep0" candidate of one match (:729-790), one lane per match.
	// pk = len | len_test_2 << 9 | (byte at target - dist - 1) << 18, rel as in DpRec::m_rel.
	__device__ __forceinline__ void mlr_eval(uint32_t p, uint32_t pos, uint32_t ps, const xzb_pair pr, uint32_t &pk, uint32_t &dist, uint32_t &rel) const
	{
		const uint8_t *b = buf + p;
		const uint32_t L = XZB_PAIR_LEN(pr.len), r = XZB_PAIR_LEN2(pr.len), mbm = XZB_PAIR_MB(pr.len);
		dist = pr.dist;
		const uint32_t avail1 = size - p;
		const uint32_t limit = xzb_min(avail1, L + 1 + nice_len);
		uint32_t lt2 = 0, mbt = 0;
		rel = 0;
		if (L + 1 < limit) lt2 = xzb_min(L + 1 + r, limit) - (L + 1);
		if (lt2 >= 2) {
			const uint32_t psn2 = (pos + L + 1) & pos_mask;
			rel = dist_len_price(dist, L, ps) + literal_price_matched_lane(pos + L, b[L - 1], mbm, b[L])
					+ S.pb[4][psn2][1].x + len_price(1, lt2, psn2);
			mbt = *(b + L + 1 + lt2 - dist - 1);
		} else {
			lt2 = 0;
		}
		pk = L | (lt2 << 9) | (mbt << 18);
	}

	// ---- lzma_lzma_encode for one LZMA2 chunk (lzma_encoder.c:266-436) ----
	__device__ void encode_chunk(uint32_t limit)
	{
		if (!is_initialized) {
			if (read_pos != size) {
				mf_skip(1);
				read_ahead = 0;
				WSeg segs[2] = { WSeg{ PI_IS_MATCH, SEG_SINGLE, 1, 0 }, WSeg{ PI_LITERAL, SEG_TREE, 8, buf[0] } };
				encode_segments(segs, 2, 0, 0);
				++uncomp_size;
			}
			is_initialized = 1;
		}
		for (;;) {
			if (read_pos - read_ahead >= limit || rc_out_pos + (rc_cache_size + 4) >= XZB_LZMA2_CHUNK_MAX - XZB_LOOP_INPUT_MAX) break;
			if (read_pos >= size) { if (read_ahead == 0) break; }
			if (mf_stalled) break;
			uint32_t len, back;
			optimum_normal(&back, &len, uncomp_size);
			if (mf_stalled) break;
			if (trace != nullptr && lane == 0 && trace_n < trace_cap) { trace[3 * trace_n] = uncomp_size; trace[3 * trace_n + 1] = back; trace[3 * trace_n + 2] = len; }
			++trace_n;
			encode_symbol(back, len, uncomp_size);
			uncomp_size += len;
		}
		rc_flush();
	}
};

// Helper warp: for the segment announced by the DP warp, produce DpRec records for nodes 1, 2, ...
// (at most hr_mask + 1 ahead).  Reads probabilities / price tables / bundles, all frozen while a
// segment's DP runs; records of an abandoned segment are simply never consumed.
__device__ inline void xzb_dp_helper_main(DS &S, DpEnc &H)
{
	const uint32_t lane = H.lane;
	uint32_t my_epoch = 0;
	uint32_t ring_base = 0x80000000u;
	for (;;) {
		uint32_t e;
		while ((e = S.h_epoch) == my_epoch) { if (S.m_exit) return; __nanosleep(40); }
		__threadfence_block();
		my_epoch = e;
		const uint32_t P0 = S.h_pos0, position0 = S.h_position0;
		for (uint32_t c = 1; c < XZB_OPTS; ++c) {
			while (c > S.h_consumed + H.hr_mask && S.h_epoch == my_epoch && !S.m_exit) __nanosleep(20);
			if (S.h_epoch != my_epoch || S.m_exit) break;
			const uint32_t p = P0 + c;
			if (p >= H.size) break;
			const uint32_t pos = position0 + c;
			const uint32_t ps = pos & H.pos_mask;
			DpRec &R = S.rec[c & H.hr_mask];
			if (p - ring_base >= 32u) {  // refill the helper's own view of the match store
				__syncwarp();
				ring_base = p;
				const uint32_t need = xzb_min(p + 32, H.size);
				if (need > H.mf_done) H.mf_wait(need);
				if (H.mf_stalled) {
					if (lane == 0) { R.hdr = DP_STALL_HDR; __threadfence_block(); R.tag = ((my_epoch << 16) | c) + 1; }
					break;
				}
				const uint32_t g = p + lane;
				if (g < H.size) {
					S.mring_mh[lane] = H.g_mh[g];
					const uint4 *src = reinterpret_cast<const uint4 *>(H.g_mp + (size_t)g * 8);
					uint4 *dst = reinterpret_cast<uint4 *>(&S.mring_mp[lane][0]);
					const uint4 a = src[0], b = src[1], cc = src[2], d = src[3];
					dst[0] = a; dst[1] = b; dst[2] = cc; dst[3] = d;
				}
				__syncwarp();
			}
			const uint32_t slot = p - ring_base;
			const uint32_t h = S.mring_mh[slot];
			const uint32_t count = h & 0xFFFF, longest = h >> 16;
			const uint8_t *b = H.buf + p;
			const uint32_t cb = b[0], pbyte = b[-1];
			if (longest < H.nice_len) {
				const uint32_t lit = H.literal_price(pos, pbyte, false, 0, cb);
				if (lane == 0) R.lit_plain = lit;
				// matches 0..31: one lane each
				uint32_t L = 0, dist = 0;
				if (lane < count) {
					xzb_pair pr;
					if (count <= 8 || lane < 7) pr = S.mring_mp[slot][lane];
					else { const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(H.g_ovf + S.mring_mp[slot][7].len + (lane - 7))); pr = xzb_pair{ v.x, v.y }; }
					uint32_t pk, rel;
					H.mlr_eval(p, pos, ps, pr, pk, dist, rel);
					L = pk & 0x1FF;
					R.m_pack[lane] = pk; R.m_dist[lane] = dist; R.m_rel[lane] = rel;
				}
				// per-length table: dist of the first match that covers the length
				uint2 *PL = H.plain_of(c);
				const uint32_t c32 = xzb_min(count, 32u);
				for (uint32_t l = 2 + lane; l - lane <= longest; l += 32) {  // uniform trip count: shuffles inside
					uint32_t idx = 0;
					for (uint32_t j = 0; j + 1 < c32; ++j) { const uint32_t Lj = __shfl_sync(WFULL, L, j); if (Lj < l) idx = j + 1; }
					uint32_t di = __shfl_sync(WFULL, dist, idx & 31);
					if (count > 32 && idx == 31) {  // beyond the 32 lanes: walk the rest of the list
						for (uint32_t j = 31; j < count; ++j) {
							const xzb_pair pr = H.match_pair(p, j);
							di = pr.dist;
							if (XZB_PAIR_LEN(pr.len) >= l) break;
						}
					}
					if (l <= longest) {
						const uint32_t mbt = *(b + l - di - 1);
						PL[l - 2] = make_uint2(H.dist_len_price(di, l, ps) | (mbt << 16), di);
					}
				}
			}
			if (lane == 0) { R.hdr = h; R.bytes = cb | (pbyte << 8); }
			__syncwarp();
			__threadfence_block();
			if (lane == 0) R.tag = ((my_epoch << 16) | c) + 1;
			if (longest >= H.nice_len) break;  // the DP loop stops at this position
		}
	}
}
