// xzb_filters.cuh -- the non-LZMA2 filters of the .xz filter chain table (common/filter_encoder.c:59-182,
// filter_decoder.c:44-139): Delta and the BCJ ("simple") filters, as whole-Block transforms (host/device).
//
// In the reference every Block gets a fresh filter chain and the filters see the Block's bytes as a stream
// (delta/delta_encoder.c, delta_decoder.c, simple/simple_coder.c); each is a pure function of (position in the
// Block, bytes), so here one call transforms the whole Block in place between the host copy and the LZMA2 encoder
// (encode) or between the LZMA2 decoder and the integrity check (decode).  Bytes at the end of the Block that a BCJ
// filter cannot complete an instruction with stay as they are (simple_coder.c:173-176).
//
//   delta        delta/delta_encoder.c:20-46, delta_decoder.c:20-33   out[i] = in[i] -/+ x[i - dist], x = 0 before the Block
//   x86          simple/x86.c:22-115      sequential (prev_mask / prev_pos state)
//   powerpc      simple/powerpc.c:15-52   4-byte units, independent
//   ia64         simple/ia64.c:15-88      16-byte bundles, independent
//   arm          simple/arm.c:15-46       4-byte units, independent
//   armthumb     simple/armthumb.c:15-52  2-byte steps, a converted BL skips the next unit: sequential
//   sparc        simple/sparc.c:15-57     4-byte units, independent
//   arm64        simple/arm64.c:22-106    4-byte units, independent
//   riscv        simple/riscv.c:495-747   2-byte steps, JAL (4 bytes) and AUIPC + I-type pairs (8 bytes): sequential
// The independent ones run one unit per thread (xzb_k_filter_units), the sequential ones on one thread per Block.
#pragma once
#include "xzb_common.cuh"

#define XZB_FILTER_DELTA 0x03u
#define XZB_FILTER_X86 0x04u
#define XZB_FILTER_POWERPC 0x05u
#define XZB_FILTER_IA64 0x06u
#define XZB_FILTER_ARM 0x07u
#define XZB_FILTER_ARMTHUMB 0x08u
#define XZB_FILTER_SPARC 0x09u
#define XZB_FILTER_ARM64 0x0Au
#define XZB_FILTER_RISCV 0x0Bu
#define XZB_FILTER_LZMA2 0x21u
#define XZB_FILTERS_MAX 4u

struct XzbPreFilter {      // one non-last filter of a chain
	uint32_t id;
	uint32_t arg;          // delta: distance 1..256; BCJ: start_offset
};

XZB_HD uint32_t xzb_filter_unit(uint32_t id)   // bytes per independent unit, 0 = sequential filter
{
	switch (id) {
	case XZB_FILTER_POWERPC: case XZB_FILTER_ARM: case XZB_FILTER_SPARC: case XZB_FILTER_ARM64: return 4;
	case XZB_FILTER_IA64: return 16;
	default: return 0;
	}
}
XZB_HD uint32_t xzb_filter_alignment(uint32_t id)   // start_offset must be a multiple of this (simple_coder.c:276-278)
{
	switch (id) {
	case XZB_FILTER_X86: return 1;
	case XZB_FILTER_ARMTHUMB: case XZB_FILTER_RISCV: return 2;
	case XZB_FILTER_IA64: return 16;
	default: return 4;
	}
}
XZB_HD bool xzb_filter_known(uint32_t id) { return id >= XZB_FILTER_DELTA && id <= XZB_FILTER_RISCV; }

// One independent unit at buffer offset i (a multiple of the unit size); pc = start_offset + i.
XZB_HD void xzb_bcj_unit(uint32_t id, uint8_t *p, uint32_t pc, bool enc)
{
	if (id == XZB_FILTER_ARM) {            // BL: 24-bit word offset, pipeline offset 8
		if (p[3] != 0xEB) return;
		uint32_t src = ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
		src <<= 2;
		uint32_t dest = enc ? pc + 8 + src : src - (pc + 8);
		dest >>= 2;
		p[2] = (uint8_t)(dest >> 16); p[1] = (uint8_t)(dest >> 8); p[0] = (uint8_t)dest;
	} else if (id == XZB_FILTER_POWERPC) {  // b/bl with AA = 0, LK = 1: 6-bit opcode 18, 24-bit offset
		if ((p[0] >> 2) != 0x12 || (p[3] & 3) != 1) return;
		const uint32_t src = (((uint32_t)p[0] & 3) << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | ((uint32_t)p[3] & ~3u);
		const uint32_t dest = enc ? pc + src : src - pc;
		p[0] = (uint8_t)(0x48 | ((dest >> 24) & 3)); p[1] = (uint8_t)(dest >> 16); p[2] = (uint8_t)(dest >> 8);
		p[3] = (uint8_t)((p[3] & 3) | (dest & 0xFC));
	} else if (id == XZB_FILTER_SPARC) {    // call: 01 + 30-bit displacement, only +-4 MiB-ish targets (sign-extended 22 bits)
		if (!((p[0] == 0x40 && (p[1] & 0xC0) == 0x00) || (p[0] == 0x7F && (p[1] & 0xC0) == 0xC0))) return;
		uint32_t src = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
		src <<= 2;
		uint32_t dest = enc ? pc + src : src - pc;
		dest >>= 2;
		dest = (((0u - ((dest >> 22) & 1)) << 22) & 0x3FFFFFFFu) | (dest & 0x3FFFFFu) | 0x40000000u;
		p[0] = (uint8_t)(dest >> 24); p[1] = (uint8_t)(dest >> 16); p[2] = (uint8_t)(dest >> 8); p[3] = (uint8_t)dest;
	} else if (id == XZB_FILTER_ARM64) {
		uint32_t instr = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
		if ((instr >> 26) == 0x25) {         // BL: the whole 26-bit immediate
			uint32_t w = pc >> 2;
			if (!enc) w = 0u - w;
			instr = 0x94000000u | ((instr + w) & 0x03FFFFFFu);
		} else if ((instr & 0x9F000000u) == 0x90000000u) {   // ADRP within +-512 MiB
			const uint32_t src = ((instr >> 29) & 3) | ((instr >> 3) & 0x001FFFFCu);
			if ((src + 0x00020000u) & 0x001C0000u) return;
			instr &= 0x9000001Fu;
			uint32_t w = pc >> 12;
			if (!enc) w = 0u - w;
			const uint32_t dest = src + w;
			instr |= (dest & 3) << 29;
			instr |= (dest & 0x0003FFFCu) << 3;
			instr |= (0u - (dest & 0x00020000u)) & 0x00E00000u;
		} else {
			return;
		}
		p[0] = (uint8_t)instr; p[1] = (uint8_t)(instr >> 8); p[2] = (uint8_t)(instr >> 16); p[3] = (uint8_t)(instr >> 24);
	} else if (id == XZB_FILTER_IA64) {     // 128-bit bundle: template selects which of the three 41-bit slots hold a branch
		const uint32_t tmpl = p[0] & 0x1F;
		static const uint8_t kBranch[32] = { 0,0,0,0,0,0,0,0, 0,0,0,0,0,0,0,0, 4,4,6,6,0,0,7,7, 4,4,0,0,4,4,0,0 };
		const uint32_t m = kBranch[tmpl];
		uint32_t bit_pos = 5;
		for (uint32_t slot = 0; slot < 3; ++slot, bit_pos += 41) {
			if (((m >> slot) & 1) == 0) continue;
			const uint32_t byte_pos = bit_pos >> 3, bit_res = bit_pos & 7;
			uint64_t instruction = 0;
			for (uint32_t j = 0; j < 6; ++j) instruction += (uint64_t)p[j + byte_pos] << (8 * j);
			uint64_t norm = instruction >> bit_res;
			if (((norm >> 37) & 0xF) != 0x5 || ((norm >> 9) & 0x7) != 0) continue;
			uint32_t src = (uint32_t)((norm >> 13) & 0xFFFFF);
			src |= (uint32_t)((norm >> 36) & 1) << 20;
			src <<= 4;
			uint32_t dest = enc ? pc + src : src - pc;
			dest >>= 4;
			norm &= ~((uint64_t)0x8FFFFF << 13);
			norm |= (uint64_t)(dest & 0xFFFFF) << 13;
			norm |= (uint64_t)(dest & 0x100000) << (36 - 20);
			instruction &= (1u << bit_res) - 1;
			instruction |= norm << bit_res;
			for (uint32_t j = 0; j < 6; ++j) p[j + byte_pos] = (uint8_t)(instruction >> (8 * j));
		}
	}
}

// Sequential filters over a whole Block, in place.  Returns the bytes processed (the rest is left as it is).
XZB_HD uint32_t xzb_bcj_x86(uint8_t *buf, uint32_t size, uint32_t now_pos, bool enc)
{
	if (size < 5) return 0;
	uint32_t prev_mask = 0, prev_pos = now_pos - 5;   // x86_coder_init: prev_pos = -5, then "now_pos - prev_pos > 5" -> now_pos - 5
	const uint32_t limit = size - 5;
	uint32_t i = 0;
	while (i <= limit) {
		uint8_t b = buf[i];
		if (b != 0xE8 && b != 0xE9) { ++i; continue; }
		const uint32_t offset = now_pos + i - prev_pos;
		prev_pos = now_pos + i;
		if (offset > 5) prev_mask = 0;
		else for (uint32_t k = 0; k < offset; ++k) { prev_mask &= 0x77; prev_mask <<= 1; }
		b = buf[i + 4];
		if ((b == 0 || b == 0xFF) && (prev_mask >> 1) <= 4 && (prev_mask >> 1) != 3) {
			uint32_t src = ((uint32_t)b << 24) | ((uint32_t)buf[i + 3] << 16) | ((uint32_t)buf[i + 2] << 8) | buf[i + 1];
			uint32_t dest;
			for (;;) {
				dest = enc ? src + (now_pos + i + 5) : src - (now_pos + i + 5);
				if (prev_mask == 0) break;
				const uint32_t k = (0x32210u >> (4 * (prev_mask >> 1))) & 0xF;   // MASK_TO_BIT_NUMBER { 0, 1, 2, 2, 3 }
				b = (uint8_t)(dest >> (24 - k * 8));
				if (!(b == 0 || b == 0xFF)) break;
				src = dest ^ ((1u << (32 - k * 8)) - 1);
			}
			buf[i + 4] = (uint8_t)(~(((dest >> 24) & 1) - 1));
			buf[i + 3] = (uint8_t)(dest >> 16); buf[i + 2] = (uint8_t)(dest >> 8); buf[i + 1] = (uint8_t)dest;
			i += 5;
			prev_mask = 0;
		} else {
			++i;
			prev_mask |= 1;
			if (b == 0 || b == 0xFF) prev_mask |= 0x10;
		}
	}
	return i;
}

XZB_HD uint32_t xzb_bcj_armthumb(uint8_t *buf, uint32_t size, uint32_t now_pos, bool enc)
{
	if (size < 4) return 0;
	size -= 4;
	uint32_t i;
	for (i = 0; i <= size; i += 2) {
		if ((buf[i + 1] & 0xF8) == 0xF0 && (buf[i + 3] & 0xF8) == 0xF8) {   // BL prefix + suffix halfwords
			uint32_t src = (((uint32_t)buf[i + 1] & 7) << 19) | ((uint32_t)buf[i] << 11) | (((uint32_t)buf[i + 3] & 7) << 8) | buf[i + 2];
			src <<= 1;
			uint32_t dest = enc ? now_pos + i + 4 + src : src - (now_pos + i + 4);
			dest >>= 1;
			buf[i + 1] = (uint8_t)(0xF0 | ((dest >> 19) & 7)); buf[i] = (uint8_t)(dest >> 11);
			buf[i + 3] = (uint8_t)(0xF8 | ((dest >> 8) & 7)); buf[i + 2] = (uint8_t)dest;
			i += 2;
		}
	}
	return i;
}

// RISC-V (simple/riscv.c).  Two kinds of instructions carry pc-relative targets worth converting:
//   JAL with rd = x1 / x5 (byte 0 == 0xEF, rd bits in byte 1): the scrambled 20-bit immediate becomes the absolute
//     address, stored big endian in bytes 1..3 (:508-548 encode, :640-672 decode);
//   AUIPC (low 7 bits 0x17) followed by an I-type instruction that uses AUIPC's rd as rs1: the pair's 32-bit absolute
//     address is stored big endian in the second word and the pair is rewritten as a "special" AUIPC with rd = x2
//     that carries the second instruction's low bits (:550-604 / :674-728).  An AUIPC that already looks special
//     (rd = x0 or x2) is escaped into the ordinary form so that decoding stays unambiguous.
// The tests on the instruction words are the reference's bit expressions: a pair matches when
// ((auipc << 8) ^ (inst2 - 3)) & 0xF8003 == 0 (same register, 32-bit instruction); a special AUIPC is kept as it is
// when (auipc - 0x3117) << 18 >= (rs1 & 0x1D) as unsigned 32-bit values.
XZB_HD uint32_t xzb_rd32le(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
XZB_HD void xzb_wr32le(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
XZB_HD uint32_t xzb_rd32be(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
XZB_HD void xzb_wr32be(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

XZB_HD uint32_t xzb_bcj_riscv(uint8_t *buf, uint32_t size, uint32_t now_pos, bool enc)
{
	if (size < 8) return 0;
	size -= 8;
	uint32_t i;
	for (i = 0; i <= size; i += 2) {
		uint32_t inst = buf[i];
		if (inst == 0xEF) {                       // JAL
			const uint32_t b1 = buf[i + 1];
			if ((b1 & 0x0D) != 0) continue;       // rd is neither x1 nor x5
			const uint32_t b2 = buf[i + 2], b3 = buf[i + 3];
			const uint32_t pc = now_pos + i;
			if (enc) {
				// imm[20|10:1|11|19:12] in instruction bits 31..12 -> address, then big endian into bytes 1..3
				uint32_t addr = ((b1 & 0xF0) << 8) | ((b2 & 0x0F) << 16) | ((b2 & 0x10) << 7) | ((b2 & 0xE0) >> 4)
						| ((b3 & 0x7F) << 4) | ((b3 & 0x80) << 13);
				addr += pc;
				buf[i + 1] = (uint8_t)((b1 & 0x0F) | ((addr >> 13) & 0xF0));
				buf[i + 2] = (uint8_t)(addr >> 9);
				buf[i + 3] = (uint8_t)(addr >> 1);
			} else {
				uint32_t addr = ((b1 & 0xF0) << 13) | (b2 << 9) | (b3 << 1);
				addr -= pc;
				buf[i + 1] = (uint8_t)((b1 & 0x0F) | ((addr >> 8) & 0xF0));
				buf[i + 2] = (uint8_t)(((addr >> 16) & 0x0F) | ((addr >> 7) & 0x10) | ((addr << 4) & 0xE0));
				buf[i + 3] = (uint8_t)(((addr >> 4) & 0x7F) | ((addr >> 13) & 0x80));
			}
			i += 4 - 2;
		} else if ((inst & 0x7F) == 0x17) {       // AUIPC
			inst |= (uint32_t)buf[i + 1] << 8;
			inst |= (uint32_t)buf[i + 2] << 16;
			inst |= (uint32_t)buf[i + 3] << 24;
			uint32_t w0, w1;                      // the two words as they are written back
			if (inst & 0xE80) {                   // rd is neither x0 nor x2: an ordinary AUIPC
				const uint32_t inst2 = xzb_rd32le(buf + i + 4);
				if ((((inst << 8) ^ (inst2 - 3)) & 0xF8003u) != 0) { i += 6 - 2; continue; }   // not a pair
				uint32_t addr = inst & 0xFFFFF000u;
				if (enc) {
					addr += (inst2 >> 20) - ((inst2 >> 19) & 0x1000);   // sign-extended 12-bit immediate of inst2
					addr += now_pos + i;
					w0 = 0x17u | (2u << 7) | (inst2 << 12);
					xzb_wr32le(buf + i, w0);
					xzb_wr32be(buf + i + 4, addr);
					i += 8 - 2;
					continue;
				}
				addr += inst2 >> 20;                // decoder: un-escape an AUIPC that only looked special
				w0 = 0x17u | (2u << 7) | (inst2 << 12);
				w1 = addr;
			} else {                                // rd = x0 / x2: special form
				const uint32_t rs1 = inst >> 27;
				if ((uint32_t)((inst - 0x3117u) << 18) >= (rs1 & 0x1Du)) { i += 4 - 2; continue; }
				if (enc) {                          // encoder: escape it
					const uint32_t fake_addr = xzb_rd32le(buf + i + 4);
					w1 = (inst >> 12) | (fake_addr << 20);
					w0 = 0x17u | (rs1 << 7) | (fake_addr & 0xFFFFF000u);
				} else {                            // decoder: rebuild the pair from the absolute address
					uint32_t addr = xzb_rd32be(buf + i + 4);
					addr -= now_pos + i;
					w1 = (inst >> 12) | (addr << 20);
					w0 = 0x17u | (rs1 << 7) | ((addr + 0x800u) & 0xFFFFF000u);
				}
			}
			xzb_wr32le(buf + i, w0);
			xzb_wr32le(buf + i + 4, w1);
			i += 8 - 2;
		}
	}
	return i;
}

// Delta, sequential form for the decoder side of a unit test / the host; the kernels use the closed forms
// enc: out[i] = in[i] - in[i - d]   dec: out[i] = in[i] + out[i - d]  (bytes before the Block are 0)
XZB_HD void xzb_delta_seq(uint8_t *buf, uint32_t size, uint32_t dist, bool enc)
{
	if (enc) { for (uint32_t i = size; i-- > dist;) buf[i] = (uint8_t)(buf[i] - buf[i - dist]); }
	else { for (uint32_t i = dist; i < size; ++i) buf[i] = (uint8_t)(buf[i] + buf[i - dist]); }
}

// Whole-Block application of one filter on one thread (host side of the tests; the device kernels split the work).
XZB_HD void xzb_filter_apply_seq(const XzbPreFilter f, uint8_t *buf, uint32_t size, bool enc)
{
	if (f.id == XZB_FILTER_DELTA) { xzb_delta_seq(buf, size, f.arg, enc); return; }
	if (f.id == XZB_FILTER_X86) { xzb_bcj_x86(buf, size, f.arg, enc); return; }
	if (f.id == XZB_FILTER_ARMTHUMB) { xzb_bcj_armthumb(buf, size, f.arg, enc); return; }
	if (f.id == XZB_FILTER_RISCV) { xzb_bcj_riscv(buf, size, f.arg, enc); return; }
	const uint32_t u = xzb_filter_unit(f.id);
	if (u == 0) return;
	for (uint32_t i = 0; i + u <= size; i += u) xzb_bcj_unit(f.id, buf + i, f.arg + i, enc);
}
