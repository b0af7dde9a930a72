/*
 * oracle/ref_shim.c -- thin harness around the UNMODIFIED reference liblzma
 * (oracle/_ref/liblzma_ref.so).  TEST INFRASTRUCTURE ONLY.  Compiled in the build
 * container against the reference's public headers (-I/root/reference/src/liblzma/api);
 * only the resulting oracle/_ref/libref_shim.so travels to the GPU box.
 * Call pattern = doc/examples/04_compress_easy_mt.c with in-memory buffers
 * and lzma_mt.block_size set explicitly.
 */
#include <lzma.h>
#include <string.h>
#include <stdlib.h>

/* returns lzma_ret (LZMA_OK on success); *out_size = stream size */
int ref_encode_mt(const uint8_t *in, size_t in_size, uint32_t preset, uint64_t block_size,
		uint32_t check, uint32_t threads, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_mt mt;
	memset(&mt, 0, sizeof(mt));
	mt.flags = 0;
	mt.threads = threads ? threads : lzma_cputhreads();
	if (mt.threads == 0) mt.threads = 1;
	mt.block_size = block_size;
	mt.timeout = 0;
	mt.preset = preset;
	mt.filters = NULL;
	mt.check = (lzma_check)check;
	lzma_ret ret = lzma_stream_encoder_mt(&strm, &mt);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* Same but with explicit LZMA2 options instead of a preset. */
int ref_encode_mt_opts(const uint8_t *in, size_t in_size, uint32_t dict_size, uint32_t lc, uint32_t lp,
		uint32_t pb, uint32_t mode, uint32_t nice_len, uint32_t mf, uint32_t depth,
		uint64_t block_size, uint32_t check, uint32_t threads, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_options_lzma o;
	memset(&o, 0, sizeof(o));
	o.dict_size = dict_size; o.lc = lc; o.lp = lp; o.pb = pb; o.mode = (lzma_mode)mode;
	o.nice_len = nice_len; o.mf = (lzma_match_finder)mf; o.depth = depth;
	lzma_filter filters[2] = { { LZMA_FILTER_LZMA2, &o }, { LZMA_VLI_UNKNOWN, NULL } };
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_mt mt;
	memset(&mt, 0, sizeof(mt));
	mt.threads = threads ? threads : lzma_cputhreads();
	if (mt.threads == 0) mt.threads = 1;
	mt.block_size = block_size;
	mt.filters = filters;
	mt.check = (lzma_check)check;
	lzma_ret ret = lzma_stream_encoder_mt(&strm, &mt);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* lzma_stream_decoder (single-threaded), flags as given; returns lzma_ret mapped so
 * that LZMA_STREAM_END -> LZMA_OK. */
int ref_decode(const uint8_t *in, size_t in_size, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_ret ret = lzma_stream_decoder(&strm, UINT64_MAX, 0);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* lzma_stream_decoder with caller-chosen flags (e.g. LZMA_CONCATENATED). */
int ref_decode_flags(const uint8_t *in, size_t in_size, uint32_t flags, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_ret ret = lzma_stream_decoder(&strm, UINT64_MAX, flags);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* lzma_stream_decoder driven like src/xz/coder.c:1226-1352 does: every return code other than LZMA_OK is
 * recorded (LZMA_NO_CHECK / LZMA_UNSUPPORTED_CHECK / LZMA_GET_CHECK are followed by lzma_get_check()
 * and the loop goes on).  codes[i] = lzma_ret | check << 8. */
int ref_decode_trace_memlimit(const uint8_t *in, size_t in_size, uint32_t flags, uint64_t memlimit, uint8_t *out, size_t out_cap, size_t *out_size,
		uint32_t *codes, uint32_t codes_cap, uint32_t *n_codes, uint64_t *memusage_seen);
int ref_decode_trace(const uint8_t *in, size_t in_size, uint32_t flags, uint8_t *out, size_t out_cap, size_t *out_size,
		uint32_t *codes, uint32_t codes_cap, uint32_t *n_codes)
{
	uint64_t mu = 0;
	return ref_decode_trace_memlimit(in, in_size, flags, UINT64_MAX, out, out_cap, out_size, codes, codes_cap, n_codes, &mu);
}

/* Same with a memory limit: LZMA_MEMLIMIT_ERROR is recorded, *memusage_seen = lzma_memusage() at that point, the
 * limit is raised to exactly that with lzma_memlimit_set() (its return value is recorded as code | 0x8000)
 * and decoding goes on (src/xz/coder.c:1292-1316 shows the pattern). */
int ref_decode_trace_memlimit(const uint8_t *in, size_t in_size, uint32_t flags, uint64_t memlimit, uint8_t *out, size_t out_cap, size_t *out_size,
		uint32_t *codes, uint32_t codes_cap, uint32_t *n_codes, uint64_t *memusage_seen)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	*n_codes = 0; *out_size = 0; *memusage_seen = 0;
	lzma_ret ret = lzma_stream_decoder(&strm, memlimit, flags);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	for (;;) {
		ret = lzma_code(&strm, LZMA_FINISH);
		if (ret == LZMA_OK) continue;
		if (*n_codes < codes_cap) codes[(*n_codes)++] = (uint32_t)ret | ((uint32_t)lzma_get_check(&strm) << 8);
		if (ret == LZMA_NO_CHECK || ret == LZMA_UNSUPPORTED_CHECK || ret == LZMA_GET_CHECK) continue;
		if (ret == LZMA_MEMLIMIT_ERROR && *memusage_seen == 0) {
			*memusage_seen = lzma_memusage(&strm);
			const lzma_ret too_low = lzma_memlimit_set(&strm, *memusage_seen - 1);
			const lzma_ret ok = lzma_memlimit_set(&strm, *memusage_seen);
			if (*n_codes < codes_cap) codes[(*n_codes)++] = 0x8000u | (uint32_t)too_low | ((uint32_t)ok << 8);
			if (ok == LZMA_OK) continue;
		}
		break;
	}
	if (*memusage_seen == 0) *memusage_seen = lzma_memusage(&strm);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return (int)ret;
}

/* lzma_stream_decoder_mt with all host threads. */
int ref_decode_mt(const uint8_t *in, size_t in_size, uint32_t threads, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_mt mt;
	memset(&mt, 0, sizeof(mt));
	mt.threads = threads ? threads : lzma_cputhreads();
	if (mt.threads == 0) mt.threads = 1;
	mt.memlimit_threading = UINT64_MAX;
	mt.memlimit_stop = UINT64_MAX;
	lzma_ret ret = lzma_stream_decoder_mt(&strm, &mt);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* One-shot buffer API of the reference. */
int ref_easy_buffer_encode(const uint8_t *in, size_t in_size, uint32_t preset, uint32_t check, uint8_t *out, size_t out_cap, size_t *out_size)
{
	size_t pos = 0;
	lzma_ret ret = lzma_easy_buffer_encode(preset, (lzma_check)check, NULL, in, in_size, out, &pos, out_cap);
	*out_size = pos;
	return (int)ret;
}
size_t ref_stream_buffer_bound(size_t n) { return lzma_stream_buffer_bound(n); }
int ref_stream_buffer_decode(const uint8_t *in, size_t in_size, uint32_t flags, uint8_t *out, size_t out_cap, size_t *in_used, size_t *out_size)
{
	uint64_t memlimit = UINT64_MAX;
	size_t ip = 0, op = 0;
	lzma_ret ret = lzma_stream_buffer_decode(&memlimit, flags, NULL, in, &ip, in_size, out, &op, out_cap);
	*in_used = ip; *out_size = op;
	return (int)ret;
}

/* lzma_block_buffer_encode of the reference with the preset's LZMA2 filter. */
int ref_block_buffer_encode(const uint8_t *in, size_t in_size, uint32_t preset, uint32_t check, uint8_t *out, size_t out_cap, size_t *out_size,
		uint32_t *header_size, uint64_t *compressed_size, uint64_t *uncompressed_size, uint8_t raw_check[64])
{
	lzma_options_lzma opt;
	if (lzma_lzma_preset(&opt, preset)) return (int)LZMA_OPTIONS_ERROR;
	lzma_filter f[2] = { { LZMA_FILTER_LZMA2, &opt }, { LZMA_VLI_UNKNOWN, NULL } };
	lzma_block b;
	memset(&b, 0, sizeof(b));
	b.version = 0; b.check = (lzma_check)check; b.filters = f;
	size_t pos = 0;
	lzma_ret ret = lzma_block_buffer_encode(&b, NULL, in, in_size, out, &pos, out_cap);
	*out_size = pos;
	*header_size = b.header_size; *compressed_size = b.compressed_size; *uncompressed_size = b.uncompressed_size;
	memcpy(raw_check, b.raw_check, 64);
	return (int)ret;
}

uint32_t ref_cputhreads(void) { return lzma_cputhreads(); }
uint32_t ref_crc32(const uint8_t *b, size_t n, uint32_t c) { return lzma_crc32(b, n, c); }
uint64_t ref_crc64(const uint8_t *b, size_t n, uint64_t c) { return lzma_crc64(b, n, c); }
const char *ref_version(void) { return lzma_version_string(); }

/* ---- filter chains (Delta / BCJ in front of LZMA2), still the unmodified reference ---- */
static void chain_fill(lzma_filter *f, lzma_options_delta *od, lzma_options_bcj *ob, lzma_options_lzma *ol,
		const uint32_t *ids, const uint32_t *args, uint32_t n_pre, uint32_t preset)
{
	uint32_t k = 0;
	for (; k < n_pre; ++k) {
		f[k].id = ids[k];
		if (ids[k] == LZMA_FILTER_DELTA) {
			memset(&od[k], 0, sizeof(od[k]));
			od[k].type = LZMA_DELTA_TYPE_BYTE; od[k].dist = args[k];
			f[k].options = &od[k];
		} else {
			memset(&ob[k], 0, sizeof(ob[k]));
			ob[k].start_offset = args[k];
			f[k].options = args[k] ? &ob[k] : NULL;
		}
	}
	lzma_lzma_preset(ol, preset);
	f[k].id = LZMA_FILTER_LZMA2; f[k].options = ol;
	f[k + 1].id = LZMA_VLI_UNKNOWN; f[k + 1].options = NULL;
}

/* The bytes the LZMA2 encoder sees (enc = 1) after the n_pre filters, or what the inverse filters make of `in`
 * (enc = 0): raw encoder with the chain + raw decoder with LZMA2 alone, and the other way round. */
int ref_filter_apply(const uint32_t *ids, const uint32_t *args, uint32_t n_pre, int enc, const uint8_t *in, size_t n, uint8_t *out)
{
	lzma_filter full[LZMA_FILTERS_MAX + 1], last[2];
	lzma_options_delta od[4]; lzma_options_bcj ob[4]; lzma_options_lzma ol, ol2;
	chain_fill(full, od, ob, &ol, ids, args, n_pre, 0);
	lzma_lzma_preset(&ol2, 0);
	last[0].id = LZMA_FILTER_LZMA2; last[0].options = &ol2;
	last[1].id = LZMA_VLI_UNKNOWN; last[1].options = NULL;
	const size_t cap = n + n / 3 + 65536;
	uint8_t *tmp = malloc(cap);
	if (tmp == NULL) return LZMA_MEM_ERROR;
	size_t tp = 0, ip = 0, op = 0;
	lzma_ret r = lzma_raw_buffer_encode(enc ? full : last, NULL, in, n, tmp, &tp, cap);
	if (r == LZMA_OK) r = lzma_raw_buffer_decode(enc ? last : full, NULL, tmp, &ip, tp, out, &op, n);
	free(tmp);
	if (r == LZMA_OK && op != n) r = LZMA_DATA_ERROR;
	return (int)r;
}

/* lzma_stream_encoder_mt with a filter chain: ids/args = the filters in front of LZMA2 (preset gives its options) */
int ref_encode_mt_chain(const uint8_t *in, size_t in_size, const uint32_t *ids, const uint32_t *args, uint32_t n_pre, uint32_t preset,
		uint64_t block_size, uint32_t check, uint32_t threads, uint8_t *out, size_t out_cap, size_t *out_size)
{
	lzma_filter full[LZMA_FILTERS_MAX + 1];
	lzma_options_delta od[4]; lzma_options_bcj ob[4]; lzma_options_lzma ol;
	chain_fill(full, od, ob, &ol, ids, args, n_pre, preset);
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_mt mt;
	memset(&mt, 0, sizeof(mt));
	mt.threads = threads ? threads : lzma_cputhreads();
	if (mt.threads == 0) mt.threads = 1;
	mt.block_size = block_size;
	mt.filters = full;
	mt.check = (lzma_check)check;
	lzma_ret ret = lzma_stream_encoder_mt(&strm, &mt);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	do { ret = lzma_code(&strm, LZMA_FINISH); } while (ret == LZMA_OK);
	*out_size = strm.total_out;
	lzma_end(&strm);
	return ret == LZMA_STREAM_END ? LZMA_OK : (int)ret;
}

/* ref_decode_trace through lzma_stream_decoder_mt (threads > 1, no memory limits): the threaded decoder's own
 * lzma_code() sequences (stream_decoder_mt.c), recorded next to the single-threaded ones. */
int ref_decode_trace_mt(const uint8_t *in, size_t in_size, uint32_t flags, uint32_t threads, uint8_t *out, size_t out_cap, size_t *out_size,
		uint32_t *codes, uint32_t codes_cap, uint32_t *n_codes)
{
	lzma_stream strm = LZMA_STREAM_INIT;
	lzma_mt mt;
	memset(&mt, 0, sizeof(mt));
	mt.flags = flags; mt.threads = threads; mt.timeout = 0;
	mt.memlimit_threading = UINT64_MAX; mt.memlimit_stop = UINT64_MAX;
	*n_codes = 0; *out_size = 0;
	lzma_ret ret = lzma_stream_decoder_mt(&strm, &mt);
	if (ret != LZMA_OK) return (int)ret;
	strm.next_in = in; strm.avail_in = in_size;
	strm.next_out = out; strm.avail_out = out_cap;
	for (;;) {
		ret = lzma_code(&strm, LZMA_FINISH);
		if (ret == LZMA_OK) continue;
		if (*n_codes < codes_cap) codes[(*n_codes)++] = (uint32_t)ret | ((uint32_t)lzma_get_check(&strm) << 8);
		if (ret == LZMA_NO_CHECK || ret == LZMA_UNSUPPORTED_CHECK || ret == LZMA_GET_CHECK) continue;
		break;
	}
	*out_size = strm.total_out;
	lzma_end(&strm);
	return (int)ret;
}
