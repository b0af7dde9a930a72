// xzb_frame.cuh -- .xz container framing around the LZMA2 payload of one block, plus the
// Stream Header / Index / Stream Footer.  Host/device.  Reference: common/block_header_encoder.c,
// common/block_encoder.c:104-133, common/block_buffer_encoder.c:27-162, common/vli_encoder.c,
// common/index_encoder.c:43-165, common/stream_flags_encoder.c:29-85.
#pragma once
#include "xzb_common.cuh"

XZB_HD uint32_t xzb_vli_put(uint8_t *out, uint64_t v)  // vli_encoder.c:16-69 (single call)
{
	uint32_t n = 0;
	while (v >= 0x80) { out[n++] = (uint8_t)v | 0x80; v >>= 7; }
	out[n++] = (uint8_t)v;
	return n;
}
XZB_HD uint32_t xzb_vli_size(uint64_t v) { uint32_t n = 0; do { v >>= 7; ++n; } while (v != 0); return n; }  // vli_size.c:15-30

XZB_HD uint32_t xzb_crc32_bytes(const uint32_t *table, const uint8_t *buf, uint32_t size, uint32_t crc)
{
	crc = ~crc;
	for (uint32_t i = 0; i < size; ++i) crc = table[(crc ^ buf[i]) & 0xFF] ^ (crc >> 8);
	return ~crc;
}

XZB_HD uint32_t xzb_check_size(uint32_t check) { return check == 0 ? 0 : check == 1 ? 4 : check == 4 ? 8 : check == 10 ? 32 : 0xFFFFFFFFu; }  // CRC32, CRC64, SHA-256

// lzma2_bound + lzma_block_buffer_bound64, block_buffer_encoder.c:27-71
XZB_HD uint64_t xzb_lzma2_bound(uint64_t u) { return u + ((u + XZB_LZMA2_CHUNK_MAX - 1) / XZB_LZMA2_CHUNK_MAX) * 3 + 1; }
XZB_HD uint64_t xzbi_block_bound(uint64_t u) { return 92 + ((xzb_lzma2_bound(u) + 3) & ~(uint64_t)3); }

// lzma_block_header_size :16-68 with both sizes present: LZMA2 last, ff_len bytes of Filter Flags of the filters before it
XZB_HD uint32_t xzb_block_header_size(uint64_t comp, uint64_t uncomp, uint32_t ff_len = 0) { return (6 + xzb_vli_size(comp) + xzb_vli_size(uncomp) + ff_len + 3 + 3) & ~3u; }

// lzma_block_header_encode :71-131
XZB_HD void xzb_block_header_encode(const uint32_t *crc32_table, uint8_t *out, uint32_t header_size, uint64_t comp, uint64_t uncomp, uint8_t dict_prop,
		const uint8_t *ff = nullptr, uint32_t ff_len = 0, uint32_t n_pre = 0)
{
	const uint32_t out_size = header_size - 4;
	out[0] = (uint8_t)(out_size / 4);
	out[1] = (uint8_t)(0xC0 | n_pre);   // both sizes present, number of filters - 1
	uint32_t pos = 2;
	pos += xzb_vli_put(out + pos, comp);
	pos += xzb_vli_put(out + pos, uncomp);
	for (uint32_t i = 0; i < ff_len; ++i) out[pos++] = ff[i];   // filter_flags_encoder.c:31-56 for the Delta / BCJ filters
	out[pos++] = 0x21; out[pos++] = 0x01; out[pos++] = dict_prop;  // filter_flags_encoder.c:31-56
	while (pos < out_size) out[pos++] = 0;
	const uint32_t crc = xzb_crc32_bytes(crc32_table, out, out_size, 0);
	for (int i = 0; i < 4; ++i) out[out_size + i] = (uint8_t)(crc >> (8 * i));
}

// lzma_lzma2_props_encode, lzma/lzma2_encoder.c:375-400
XZB_HD uint8_t xzb_lzma2_dict_prop(uint32_t dict_size)
{
	uint32_t d = dict_size > 4096 ? dict_size : 4096;
	--d; d |= d >> 2; d |= d >> 3; d |= d >> 4; d |= d >> 8; d |= d >> 16;
	if (d == 0xFFFFFFFFu) return 40;
	return (uint8_t)(xzb_dist_slot(d + 1) - 24);
}

XZB_HD uint32_t xzb_stream_header(const uint32_t *crc32_table, uint8_t *out, uint32_t check)  // stream_flags_encoder.c:29-53
{
	out[0] = 0xFD; out[1] = 0x37; out[2] = 0x7A; out[3] = 0x58; out[4] = 0x5A; out[5] = 0x00;
	out[6] = 0x00; out[7] = (uint8_t)check;
	const uint32_t crc = xzb_crc32_bytes(crc32_table, out + 6, 2, 0);
	for (int i = 0; i < 4; ++i) out[8 + i] = (uint8_t)(crc >> (8 * i));
	return 12;
}
XZB_HD uint32_t xzb_stream_footer(const uint32_t *crc32_table, uint8_t *out, uint32_t check, uint64_t index_size)  // :56-85
{
	const uint32_t bs = (uint32_t)(index_size / 4 - 1);
	for (int i = 0; i < 4; ++i) out[4 + i] = (uint8_t)(bs >> (8 * i));
	out[8] = 0x00; out[9] = (uint8_t)check;
	const uint32_t crc = xzb_crc32_bytes(crc32_table, out + 4, 6, 0);
	for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(crc >> (8 * i));
	out[10] = 'Y'; out[11] = 'Z';
	return 12;
}
// index_encode, index_encoder.c:43-165; out == NULL returns the size only
XZB_HD uint64_t xzbi_index_encode(const uint32_t *crc32_table, const uint64_t *unpadded, const uint64_t *uncompressed, uint64_t count, uint8_t *out)
{
	uint64_t n = 1 + xzb_vli_size(count);
	for (uint64_t i = 0; i < count; ++i) n += xzb_vli_size(unpadded[i]) + xzb_vli_size(uncompressed[i]);
	const uint64_t padded = (n + 3) & ~(uint64_t)3;
	if (out == 0) return padded + 4;
	uint64_t pos = 0;
	out[pos++] = 0x00;
	pos += xzb_vli_put(out + pos, count);
	for (uint64_t i = 0; i < count; ++i) { pos += xzb_vli_put(out + pos, unpadded[i]); pos += xzb_vli_put(out + pos, uncompressed[i]); }
	while (pos < padded) out[pos++] = 0x00;
	uint32_t crc = 0xFFFFFFFFu;  // incremental form of xzb_crc32_bytes for 64-bit sizes
	for (uint64_t i = 0; i < pos; ++i) crc = crc32_table[(crc ^ out[i]) & 0xFF] ^ (crc >> 8);
	crc = ~crc;
	for (int i = 0; i < 4; ++i) out[pos++] = (uint8_t)(crc >> (8 * i));
	return pos;
}

// ---- per-block framing as worker_encode() does it (common/stream_encoder_mt.c:218-359) ----
struct XzbBlockResult {
	uint32_t total_size;     // bytes of the finished Block (header + data + padding + check)
	uint32_t header_size;
	uint64_t unpadded_size;  // lzma_block_unpadded_size(), block_util.c:53-77
	uint32_t fallback;       // 1 = stored with uncompressed LZMA2 chunks (block_buffer_encoder.c:87-162)
	uint32_t ret;            // XZB_OK or error
	uint32_t n_symbols, n_chunks_lzma, n_chunks_raw, pad_;
};

// `bytes` = the Check field as stored: little-endian CRC (check.c:146-165) or the 32 SHA-256 bytes
XZB_HD void xzb_put_check(uint8_t *out, uint32_t check, const uint8_t *bytes)
{
	const uint32_t n = xzb_check_size(check);
	for (uint32_t i = 0; i < n; ++i) out[i] = bytes[i];
}

// Normal path: payload already sits at out[header_size .. payload_end).  Returns false when the
// block does not fit -> caller takes the raw fallback.  Two framings:
//  oneshot == 0  worker_encode(): header + data + padding + check must fit
//                out_size = lzma_block_buffer_bound64(block_size) (stream_encoder_mt.c:225-344);
//  oneshot == 1  lzma_block_buffer_encode(): the LZMA2 data alone must fit
//                out_size = header_size + lzma2_bound(in_size) (block_buffer_encoder.c:165-210).
XZB_HD bool xzb_block_finish_normal(const uint32_t *crc32_table, uint8_t *out, uint32_t payload_end, uint32_t header_size,
		uint64_t out_size, uint32_t oneshot, uint32_t check, const uint8_t *check_bytes, uint32_t in_size, uint8_t dict_prop, XzbBlockResult *res,
		const uint8_t *ff = nullptr, uint32_t ff_len = 0, uint32_t n_pre = 0)
{
	const uint32_t csize = xzb_check_size(check);
	const uint32_t comp = payload_end - header_size;
	const uint32_t pad = (4 - (comp & 3)) & 3;
	if ((uint64_t)payload_end + (oneshot ? 0 : pad + csize) > out_size) return false;
	uint32_t pos = payload_end;
	for (uint32_t i = 0; i < pad; ++i) out[pos++] = 0;  // block_encoder.c:104-112
	xzb_put_check(out + pos, check, check_bytes);
	pos += csize;
	xzb_block_header_encode(crc32_table, out, header_size, comp, in_size, dict_prop, ff, ff_len, n_pre);
	res->total_size = pos; res->header_size = header_size;
	res->unpadded_size = (uint64_t)header_size + comp + csize;
	res->fallback = 0;
	return true;
}

// Raw fallback (the UNFILTERED input under an LZMA2-only header, block_buffer_encoder.c:87-162), work-shared by `nthreads` workers (tid in [0, nthreads)); worker 0 also writes
// header, control bytes, end marker, padding and check.  lzma_block_uncomp_encode.
XZB_HD void xzb_block_finish_raw(const uint32_t *crc32_table, const uint8_t *in, uint32_t in_size, uint8_t *out,
		uint32_t check, const uint8_t *check_bytes, XzbBlockResult *res, uint32_t tid, uint32_t nthreads)
{
	const uint64_t comp = xzb_lzma2_bound(in_size);
	const uint32_t hs = xzb_block_header_size(comp, in_size);
	const uint32_t csize = xzb_check_size(check);
	// payload copy: chunk c covers in[c*65536 ..), lands at hs + c*(65536+3) + 3
	for (uint32_t i = tid; i < in_size; i += nthreads) {
		const uint32_t c = i >> 16;
		out[hs + c * 3 + 3 + i] = in[i];
	}
	if (tid == 0) {
		xzb_block_header_encode(crc32_table, out, hs, comp, in_size, 0x00);
		const uint32_t nchunks = (in_size + XZB_LZMA2_CHUNK_MAX - 1) / XZB_LZMA2_CHUNK_MAX;
		for (uint32_t c = 0; c < nchunks; ++c) {
			const uint32_t off = c * XZB_LZMA2_CHUNK_MAX;
			const uint32_t copy = in_size - off < XZB_LZMA2_CHUNK_MAX ? in_size - off : XZB_LZMA2_CHUNK_MAX;
			uint8_t *h = out + hs + c * 3 + off;
			h[0] = c == 0 ? 0x01 : 0x02;
			h[1] = (uint8_t)((copy - 1) >> 8); h[2] = (uint8_t)((copy - 1) & 0xFF);
		}
		uint32_t pos = hs + nchunks * 3 + in_size;
		out[pos++] = 0x00;
		for (uint64_t i = comp; i & 3; ++i) out[pos++] = 0x00;
		xzb_put_check(out + pos, check, check_bytes);
		pos += csize;
		res->total_size = pos; res->header_size = hs;
		res->unpadded_size = (uint64_t)hs + comp + csize;
		res->fallback = 1;
	}
}
