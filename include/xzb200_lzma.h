/*
 * xzb200_lzma.h -- liblzma's own entry points for the LZMA2 .xz path, implemented on the GPU.
 *
 * libxzb200.so exports these symbols with liblzma's names, signatures and struct layouts
 * (ABI-compatible with <lzma.h> of XZ Utils 5.8; tests/test_api_cpu.py checks every offset
 * against the reference headers when they are present), so a program written against
 * liblzma's threaded stream coders -- src/xz/coder.c:836, 956-958, 1226 in the reference --
 * can be linked against it for this path.  Each declaration cites the reference declaration
 * it mirrors (paths relative to /root/reference/src/liblzma/).
 *
 * Scope: .xz Streams whose Blocks use the LZMA2-only filter chain with check None/CRC32/CRC64.
 * Anything else is rejected with the error liblzma itself uses for unsupported options.
 */
#ifndef XZB200_LZMA_H
#define XZB200_LZMA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char lzma_bool;           /* api/lzma/base.h:29 */
typedef uint64_t lzma_vli;                 /* api/lzma/vli.h:63 */
#define LZMA_VLI_UNKNOWN UINT64_MAX        /* vli.h:44 */
#define LZMA_FILTER_LZMA2 0x21ULL          /* lzma12.h:61 */
#define LZMA_FILTER_DELTA 0x03ULL          /* delta.h:23 */
#define LZMA_FILTER_X86 0x04ULL            /* bcj.h:23-53 */
#define LZMA_FILTER_POWERPC 0x05ULL
#define LZMA_FILTER_IA64 0x06ULL
#define LZMA_FILTER_ARM 0x07ULL
#define LZMA_FILTER_ARMTHUMB 0x08ULL
#define LZMA_FILTER_SPARC 0x09ULL
#define LZMA_FILTER_ARM64 0x0AULL
#define LZMA_FILTER_RISCV 0x0BULL
#define LZMA_PRESET_DEFAULT 6u             /* container.h:31 */
#define LZMA_PRESET_EXTREME (1u << 31)     /* container.h:61 */

typedef enum { LZMA_RESERVED_ENUM = 0 } lzma_reserved_enum;   /* base.h:42-44 */

typedef enum {                             /* base.h:55-271 */
	LZMA_OK = 0, LZMA_STREAM_END = 1, LZMA_NO_CHECK = 2, LZMA_UNSUPPORTED_CHECK = 3, LZMA_GET_CHECK = 4,
	LZMA_MEM_ERROR = 5, LZMA_MEMLIMIT_ERROR = 6, LZMA_FORMAT_ERROR = 7, LZMA_OPTIONS_ERROR = 8,
	LZMA_DATA_ERROR = 9, LZMA_BUF_ERROR = 10, LZMA_PROG_ERROR = 11, LZMA_SEEK_NEEDED = 12
} lzma_ret;

typedef enum {                             /* base.h:284-379 */
	LZMA_RUN = 0, LZMA_SYNC_FLUSH = 1, LZMA_FULL_FLUSH = 2, LZMA_FULL_BARRIER = 4, LZMA_FINISH = 3
} lzma_action;

typedef enum {                             /* check.h:25-66 */
	LZMA_CHECK_NONE = 0, LZMA_CHECK_CRC32 = 1, LZMA_CHECK_CRC64 = 4, LZMA_CHECK_SHA256 = 10
} lzma_check;

typedef enum { LZMA_MF_HC3 = 0x03, LZMA_MF_HC4 = 0x04, LZMA_MF_BT2 = 0x12, LZMA_MF_BT3 = 0x13, LZMA_MF_BT4 = 0x14 } lzma_match_finder; /* lzma12.h:58-110 */
typedef enum { LZMA_MODE_FAST = 1, LZMA_MODE_NORMAL = 2 } lzma_mode;  /* lzma12.h:128-138 */

typedef struct {                           /* base.h:406-470 */
	void *(*alloc)(void *opaque, size_t nmemb, size_t size);
	void (*free)(void *opaque, void *ptr);
	void *opaque;
} lzma_allocator;

typedef struct lzma_internal_s lzma_internal;

typedef struct {                           /* base.h:521-589 */
	const uint8_t *next_in;
	size_t avail_in;
	uint64_t total_in;
	uint8_t *next_out;
	size_t avail_out;
	uint64_t total_out;
	const lzma_allocator *allocator;
	lzma_internal *internal;
	void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
	uint64_t seek_pos;
	uint64_t reserved_int2;
	size_t reserved_int3, reserved_int4;
	lzma_reserved_enum reserved_enum1, reserved_enum2;
} lzma_stream;

#define LZMA_STREAM_INIT { NULL, 0, 0, NULL, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0, 0, 0, 0, LZMA_RESERVED_ENUM, LZMA_RESERVED_ENUM }  /* base.h:610-613 */

typedef struct {                           /* filter.h:41-63 */
	lzma_vli id;
	void *options;
} lzma_filter;

typedef enum { LZMA_DELTA_TYPE_BYTE = 0 } lzma_delta_type;   /* delta.h:33-35 */
typedef struct {                           /* delta.h:43-96 */
	lzma_delta_type type;
	uint32_t dist;                         /* 1..256 */
	uint32_t reserved_int1, reserved_int2, reserved_int3, reserved_int4;
	void *reserved_ptr1, *reserved_ptr2;
} lzma_options_delta;
typedef struct { uint32_t start_offset; } lzma_options_bcj;   /* bcj.h:81-101 */

typedef struct {                           /* lzma12.h:216-525 */
	uint32_t dict_size;
	const uint8_t *preset_dict;
	uint32_t preset_dict_size;
	uint32_t lc, lp, pb;
	lzma_mode mode;
	uint32_t nice_len;
	lzma_match_finder mf;
	uint32_t depth;
	uint32_t ext_flags, ext_size_low, ext_size_high;
	uint32_t reserved_int4, reserved_int5, reserved_int6, reserved_int7, reserved_int8;
	lzma_reserved_enum reserved_enum1, reserved_enum2, reserved_enum3, reserved_enum4;
	void *reserved_ptr1, *reserved_ptr2;
} lzma_options_lzma;

typedef struct {                           /* container.h:64-256 */
	uint32_t flags;
	uint32_t threads;
	uint64_t block_size;
	uint32_t timeout;
	uint32_t preset;
	const lzma_filter *filters;
	lzma_check check;
	lzma_reserved_enum reserved_enum1, reserved_enum2, reserved_enum3;
	uint32_t reserved_int1, reserved_int2, reserved_int3, reserved_int4;
	uint64_t memlimit_threading, memlimit_stop;
	uint64_t reserved_int7, reserved_int8;
	void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
} lzma_mt;

/* decoder flags, container.h:631-718 */
#define LZMA_TELL_NO_CHECK 0x01u
#define LZMA_TELL_UNSUPPORTED_CHECK 0x02u
#define LZMA_TELL_ANY_CHECK 0x04u
#define LZMA_CONCATENATED 0x08u
#define LZMA_IGNORE_CHECK 0x10u
#define LZMA_FAIL_FAST 0x20u

/* lzma/lzma_encoder_presets.c:16-63 (declared lzma12.h:560-561) */
lzma_bool lzma_lzma_preset(lzma_options_lzma *options, uint32_t preset);

/* common/stream_encoder_mt.c:1196-1208 (declared container.h:432-434): Blocks are encoded on the
 * GPU in batches instead of by worker threads; lzma_mt.threads is validated but otherwise unused. */
lzma_ret lzma_stream_encoder_mt(lzma_stream *strm, const lzma_mt *options);
/* common/stream_encoder_mt.c:1227-1278 (container.h:411-412): host + device bytes one batch needs */
uint64_t lzma_stream_encoder_mt_memusage(const lzma_mt *options);
/* common/filter_encoder.c:262-283 (container.h:468-469) */
uint64_t lzma_mt_block_size(const lzma_filter *filters);

/* common/stream_decoder.c:460-469 (container.h:742-744) and stream_decoder_mt.c:1995-2008
 * (container.h:774-776): both run the same GPU batch decoder. */
lzma_ret lzma_stream_decoder(lzma_stream *strm, uint64_t memlimit, uint32_t flags);
lzma_ret lzma_stream_decoder_mt(lzma_stream *strm, const lzma_mt *options);

/* common/common.c:203-376 and :379-389 (base.h:636, 653) -- same argument, sequence and
 * LZMA_BUF_ERROR rules as the reference. */
lzma_ret lzma_code(lzma_stream *strm, lzma_action action);
void lzma_end(lzma_stream *strm);

/* common/common.c:436-476 (base.h:691-734): decoder memory accounting with the REFERENCE's figures
 * (dictionary + 66200 bytes per LZMA2 Block on LP64) so that memlimit callers behave the same:
 * lzma_code returns LZMA_MEMLIMIT_ERROR (recoverable) when a Block of the next Stream needs more than the
 * limit; raise it with lzma_memlimit_set() and call lzma_code() again. */
uint64_t lzma_memusage(const lzma_stream *strm);
uint64_t lzma_memlimit_get(const lzma_stream *strm);
lzma_ret lzma_memlimit_set(lzma_stream *strm, uint64_t memlimit);

/* common/common.c:422-433 (check.h:149-150): Check ID of the Stream being decoded, valid after
 * LZMA_NO_CHECK / LZMA_UNSUPPORTED_CHECK / LZMA_GET_CHECK or any later lzma_code() return. */
lzma_check lzma_get_check(const lzma_stream *strm);

/* common/common.c:406-419 (base.h:672-673) */
void lzma_get_progress(lzma_stream *strm, uint64_t *progress_in, uint64_t *progress_out);
/* common/filter_encoder.c:210-241 -> stream_encoder_mt_update (stream_encoder_mt.c:914-950): new LZMA2 options for the Blocks that follow */
lzma_ret lzma_filters_update(lzma_stream *strm, const lzma_filter *filters);

typedef struct {                           /* block.h:28-303 */
	uint32_t version;
	uint32_t header_size;
	lzma_check check;
	lzma_vli compressed_size;
	lzma_vli uncompressed_size;
	lzma_filter *filters;
	uint8_t raw_check[64];
	void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3;
	uint32_t reserved_int1, reserved_int2;
	lzma_vli reserved_int3, reserved_int4, reserved_int5, reserved_int6, reserved_int7, reserved_int8;
	lzma_reserved_enum reserved_enum1, reserved_enum2, reserved_enum3, reserved_enum4;
	lzma_bool ignore_check;
	lzma_bool reserved_bool2, reserved_bool3, reserved_bool4, reserved_bool5, reserved_bool6, reserved_bool7, reserved_bool8;
} lzma_block;

/* common/block_buffer_encoder.c:213-325 (block.h:586-590): one Block (header with both sizes sized from
 * lzma2_bound(in_size), LZMA2 data or the uncompressed-chunk fallback, padding, check), no Stream framing.
 * Sets block->header_size, compressed_size, uncompressed_size and raw_check.  filters = {LZMA2, end}. */
lzma_ret lzma_block_buffer_encode(lzma_block *block, const lzma_allocator *allocator,
		const uint8_t *in, size_t in_size, uint8_t *out, size_t *out_pos, size_t out_size);

/* One-shot buffer API: common/stream_buffer_encoder.c:17-140, common/easy_buffer_encoder.c:16-27,
 * common/stream_buffer_decoder.c:14-92.  Encoder: filters must be {LZMA2, end}; in_size <= 1 GiB
 * (GPU path limit, LZMA_OPTIONS_ERROR above it).  `allocator` is accepted and unused (all coder
 * state lives in HBM). */
size_t lzma_stream_buffer_bound(size_t uncompressed_size);
lzma_ret lzma_stream_buffer_encode(lzma_filter *filters, lzma_check check, const lzma_allocator *allocator,
		const uint8_t *in, size_t in_size, uint8_t *out, size_t *out_pos, size_t out_size);
lzma_ret lzma_easy_buffer_encode(uint32_t preset, lzma_check check, const lzma_allocator *allocator,
		const uint8_t *in, size_t in_size, uint8_t *out, size_t *out_pos, size_t out_size);
lzma_ret lzma_stream_buffer_decode(uint64_t *memlimit, uint32_t flags, const lzma_allocator *allocator,
		const uint8_t *in, size_t *in_pos, size_t in_size, uint8_t *out, size_t *out_pos, size_t out_size);

/* common/block_buffer_encoder.c:74-84, check/check.c:18-39, :41-58 */
size_t lzma_block_buffer_bound(size_t uncompressed_size);
lzma_bool lzma_check_is_supported(lzma_check check);
uint32_t lzma_check_size(lzma_check check);

#ifdef __cplusplus
}
#endif
#endif
