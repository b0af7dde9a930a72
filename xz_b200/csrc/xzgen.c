/*
 * xzgen.c -- deterministic synthetic inputs for tests and bench.
 *   'T' synthetic text : the reference's own text generator scaled to any size
 *                        (tests/create_compress_files.c:110-152: 69 Lorem words, first
 *                        paragraph verbatim, then LCG n = 101771*n + 71777, seed 29).
 *   'E' enwik-style    : Zipf(s=1, integer weights) over a 65536-word synthetic vocabulary,
 *                        sentences wrapped in <page> markup; SplitMix64 seed 0x9E3779B97F4A7C15.
 *   'R' urandom-style  : SplitMix64 seed 0x5EED5EED5EED5EED, 8 bytes per step (seekable).
 *   'L' reference LCG  : tests/create_compress_files.c:93-106 (seed 5), 4 bytes per step.
 * Integer arithmetic only, so every host produces identical bytes.
 * Built as xz_b200/libxzgen.so (plain C, no CUDA); input synthesis, not part of the codec.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s)
{
	uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

typedef struct { uint8_t *out; size_t size, pos; uint64_t skip; } sink_t;
static inline int sink_full(const sink_t *s) { return s->pos >= s->size; }
static inline void put(sink_t *s, const char *p, size_t n)
{
	while (n--) {
		if (s->skip) { --s->skip; ++p; continue; }
		if (s->pos < s->size) s->out[s->pos++] = (uint8_t)*p;
		++p;
	}
}

static void gen_text(sink_t *s)
{
	static const char *lorem[] = {
		"Lorem", "ipsum", "dolor", "sit", "amet,", "consectetur",
		"adipisicing", "elit,", "sed", "do", "eiusmod", "tempor",
		"incididunt", "ut", "labore", "et", "dolore", "magna",
		"aliqua.", "Ut", "enim", "ad", "minim", "veniam,", "quis",
		"nostrud", "exercitation", "ullamco", "laboris", "nisi",
		"ut", "aliquip", "ex", "ea", "commodo", "consequat.",
		"Duis", "aute", "irure", "dolor", "in", "reprehenderit",
		"in", "voluptate", "velit", "esse", "cillum", "dolore",
		"eu", "fugiat", "nulla", "pariatur.", "Excepteur", "sint",
		"occaecat", "cupidatat", "non", "proident,", "sunt", "in",
		"culpa", "qui", "officia", "deserunt", "mollit", "anim",
		"id", "est", "laborum."
	};
	const size_t nw = sizeof(lorem) / sizeof(lorem[0]);
	for (size_t w = 0; w < nw; ++w) {
		put(s, lorem[w], strlen(lorem[w])); put(s, " ", 1);
		if (w % 7 == 6) put(s, "\n", 1);
	}
	uint32_t n = 29;
	while (!sink_full(s)) {
		put(s, "\n\n", 2);
		for (size_t w = 0; w < nw; ++w) {
			n = 101771u * n + 71777u;
			const char *word = lorem[n % nw];
			put(s, word, strlen(word)); put(s, " ", 1);
			if (w % 7 == 6) put(s, "\n", 1);
		}
	}
}

#define VOCAB 65536
static void gen_enwik(sink_t *s)
{
	/* English unigram frequencies, per 10000 (a..z) */
	static const uint16_t freq[26] = { 817, 149, 278, 425, 1270, 223, 202, 609, 697, 15, 77, 403, 241,
		675, 751, 193, 10, 599, 633, 906, 276, 98, 236, 15, 197, 7 };
	uint32_t cumf[26]; uint32_t tot = 0;
	for (int i = 0; i < 26; ++i) { tot += freq[i]; cumf[i] = tot; }
	uint64_t rs = 0x9E3779B97F4A7C15ull;
	char (*vocab)[13] = malloc((size_t)VOCAB * 13);
	uint8_t *vlen = malloc(VOCAB);
	uint64_t *cum = malloc((size_t)VOCAB * sizeof(uint64_t));
	uint64_t total = 0;
	for (uint32_t k = 0; k < VOCAB; ++k) {
		/* frequent words are short: length 2..12 grows slowly with rank */
		uint32_t len = 2 + (uint32_t)(splitmix(&rs) % 4) + (k > 64) + (k > 512) * 2 + (k > 4096) * 2 + (k > 32768) * 2;
		if (len > 12) len = 12;
		for (uint32_t i = 0; i < len; ++i) {
			const uint32_t r = (uint32_t)(splitmix(&rs) % tot);
			int c = 0; while (cumf[c] <= r) ++c;
			vocab[k][i] = (char)('a' + c);
		}
		vocab[k][len] = 0; vlen[k] = (uint8_t)len;
		total += (1ull << 36) / (k + 1);
		cum[k] = total;
	}
	size_t page_bytes = 0; int in_page = 0;
	while (!sink_full(s)) {
		if (!in_page) {
			put(s, "<page><title>", 13);
			const uint32_t tw = 1 + (uint32_t)(splitmix(&rs) % 3);
			for (uint32_t i = 0; i < tw; ++i) {
				const uint32_t k = (uint32_t)(splitmix(&rs) % VOCAB);
				char w[13]; memcpy(w, vocab[k], 13); w[0] = (char)(w[0] - 32);
				put(s, w, vlen[k]); if (i + 1 < tw) put(s, " ", 1);
			}
			put(s, "</title><text>", 14);
			in_page = 1; page_bytes = 0;
		}
		const uint32_t words = 5 + (uint32_t)(splitmix(&rs) % 21);
		for (uint32_t i = 0; i < words; ++i) {
			const uint64_t u = splitmix(&rs) % total;
			uint32_t lo = 0, hi = VOCAB - 1;
			while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (cum[mid] > u) hi = mid; else lo = mid + 1; }
			char w[13]; memcpy(w, vocab[lo], 13);
			if (i == 0) w[0] = (char)(w[0] - 32);
			put(s, w, vlen[lo]);
			put(s, i + 1 < words ? " " : ". ", i + 1 < words ? 1 : 2);
			page_bytes += vlen[lo] + 1;
		}
		if (page_bytes >= 2048) { put(s, "</text></page>\n", 15); in_page = 0; }
	}
	free(vocab); free(vlen); free(cum);
}

static void gen_random(uint8_t *out, size_t size, uint64_t offset)
{
	size_t i = 0;
	while (i < size) {
		const uint64_t idx = (offset + i) / 8;
		uint64_t st = 0x5EED5EED5EED5EEDull + idx * 0x9E3779B97F4A7C15ull;
		const uint64_t v = splitmix(&st);
		for (uint32_t b = (uint32_t)((offset + i) % 8); b < 8 && i < size; ++b, ++i) out[i] = (uint8_t)(v >> (8 * b));
	}
}

static void gen_lcg(sink_t *s)
{
	uint32_t n = 5;
	while (!sink_full(s)) {
		n = 101771u * n + 71777u;
		char b[4] = { (char)n, (char)(n >> 8), (char)(n >> 16), (char)(n >> 24) };
		put(s, b, 4);
	}
}

/* Fill out[0..size) with bytes [offset, offset+size) of the infinite stream `kind`. Returns 0 on success. */
int xzgen_fill(char kind, uint8_t *out, size_t size, uint64_t offset)
{
	sink_t s = { out, size, 0, offset };
	switch (kind) {
	case 'T': gen_text(&s); return 0;
	case 'E': gen_enwik(&s); return 0;
	case 'R': gen_random(out, size, offset); return 0;
	case 'L': gen_lcg(&s); return 0;
	default: return 1;
	}
}
