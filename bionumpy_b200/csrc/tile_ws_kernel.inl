// tile_ws_kernel.cu -- fused count (K6), warp-specialised: the dominant kernel of the hot path.
//
// One persistent CTA per SM.  The chunk streams through a ring of kNS shared-memory slots (16 KiB tile + 512 B of
// halo so the one row that crosses the tile end can finish in the slot) filled by cp.async.bulk (TMA).  Four kinds
// of warps work on different tiles of the ring at the same time; slots are handed on through mbarriers and small
// shared-memory queues only -- there is no CTA-wide barrier on the path of a tile, and no warp runs another role's code:
//
//   P  (1 lane)   tiles are dealt round-robin over the CTAs; waits for a slot to be free, starts its bulk copy -> full[slot]
//   S  (2 groups of 4 warps, alternate tiles) exact newline masks (128 B per lane), prefix over the group's four
//                 warps (the only named barrier: those four warps), sorted newline list of the tile, the tile's
//                 newline count to the workspace (one relaxed store), first newline of the halo            -> scanned[slot]
//   F  (1 warp)   line index of the tile's first byte = this CTA's previous tile + the counts every CTA published
//                 for the tiles in between (loads issued one tile ahead); entry structure of the newlines before
//                 the first row, last complete entry, first byte of the chunk; cuts the tile's rows into chunks
//                 of 32 and deals them round the row warps' queues                                          -> queue[warp]
//   R  (kRW warps) pops a chunk: ONE LANE PER ROW.  The lane validates its entry ('@', '+'), then walks its row's
//                 16-byte units in the slot: encode + validate (first and last unit masked, the others whole), a
//                 three-word window of 2-bit codes in registers, and every k-mer of a 16-base block is
//                 SHF + LOP3 + ATOMS.  No shuffles, no per-row cooperation.  The last chunk of a tile frees its slot
//                                                                                                           -> free[slot]
//
// Warp ids are in the order P < S < R < F: the SM arbiter favours the highest id among the ready warps (see kPWarp).
// Replaces io/one_line_buffer.py:44-71,139-182 + encodings/alphabet_encoding.py:34-46 + sequence/kmers.py:105-126 +
// sequence/count_encoded.py:173-177 in one pass over the chunk bytes.
// This file is the body of the kernel: tile_ws_kernel.cu includes it twice -- as namespace ws (k-mer counts, 8 ring
// slots) and as namespace wsm (minimizer counts: 4 ring slots, the freed shared memory holds the row warps'
// sliding-minimum buffers).  BNPK_WS_NAMESPACE, BNPK_WS_NS, BNPK_WS_SG, BNPK_WS_RW, BNPK_WS_MINZ and BNPK_WS_LAUNCH are set by the includer.

// Development knobs (switch parts of the row warps' work off, per-stage clocks): compiled in only with
// -DBNPK_WS_DEBUG_KNOBS (tools/dbg_sweep.sh, tools/stage_times.py); production builds fold them away.
#ifndef BNPK_WS_DBG
#ifdef BNPK_WS_DEBUG_KNOBS
#define BNPK_WS_DBG(a) ((a).start_offset)
#else
#define BNPK_WS_DBG(a) 0
#endif
#endif

namespace bnpk {
namespace BNPK_WS_NAMESPACE {

constexpr bool MINZ = BNPK_WS_MINZ != 0;

constexpr int kNS = BNPK_WS_NS;                        // ring slots (a ninth, paid for with a shorter newline list, did not help)
constexpr int kHalo = 512;
constexpr int kSlot = kTileBytes + kHalo;
constexpr int kNlCap = 1024;                    // newline positions of one tile kept in shared memory
constexpr int kWinRows = 224;                   // tiles with more newlines are walked in windows of this many rows
constexpr int kRowMax = 1024;                   // longer rows go to the deferred (one warp per segment) pass
constexpr int kMaxBins = 16384;
constexpr uint32_t kNoCross = 0xFFFFFFFFu;
constexpr int kSW = 4;                          // warps of a scan group: 4 KiB of the tile each, 128 B per lane
constexpr int kSG = BNPK_WS_SG;                        // scan groups; kNS % kSG == 0: a slot always belongs to the same group (a group must
                                                // never wait for a slot's phase u+1 before phase u completed: mbarrier parity waits
                                                // cannot tell two phases apart)
constexpr int kRW = BNPK_WS_RW;                        // row warps
constexpr int kQN = 64;                         // entries of the chunk queue
// warp ids: the SM arbiter favours the highest id among the ready warps, and a warp that spins on an mbarrier is
// always ready -- so the consumers come last: P (0) < S (1 ..) < R < F.  (With the scan warps on top, their wait for
// the next copy took most issue slots from the row warps, which are the ones that free the slots the copies need.)
constexpr int kPWarp = 0, kSWarp0 = 1, kRWarp0 = kSWarp0 + kSG * kSW, kFWarp = kRWarp0 + kRW;
constexpr int kWarps = kFWarp + 1;
constexpr int kCta = kWarps * 32;
constexpr int kFK = 6;                          // look-back loads per lane kept in flight (192 tiles)
static_assert(kSW * 4096 == kTileBytes, "scan geometry");
static_assert(kNS <= 15 && kNS % kSG == 0 && (kQN & (kQN - 1)) == 0 && kRW <= 32, "ring sizes");
static_assert(kWinRows % 32 == 0 && 4 * kWinRows + 64 <= kNlCap, "list window");

// per-slot descriptor (32-bit words)
constexpr int kDTile = 0;                       // P: tile index, 0xFFFFFFFF = end of the launch
constexpr int kDCount = 1;                      // S: newlines in the tile proper
constexpr int kDCross = 2;                      // S: first newline of the halo (slot-relative) or kNoCross
constexpr int kDRemain = 3;                     // F: chunks of the tile not finished yet
constexpr int kDBase = 4;                       // F: int64 line index of the tile's first byte
constexpr int kDescWords = 16;
constexpr int kDTIssue = 8, kDTFull = 9, kDTScanned = 10, kDTPush = 11;   // development timing (BNPK_WS_DEBUG & 16)
// chunk queue (one, shared by the row warps): entry = tag << 16 | slot << 12 | chunk, tag = (push number & 0x7FFF) + 1;
// 0 = consumed / empty
constexpr uint32_t kChunkWhole = 0xFFFu;        // the whole tile, walked in windows (more newlines than the list holds)
constexpr uint32_t kChunkEnd = 0xFFEu;
// shared memory after the histogram (bytes)
constexpr int kOffSlots = 0;
constexpr int kOffList = kOffSlots + kNS * kSlot;
constexpr int kOffDesc = kOffList + kNS * kNlCap * 2;
constexpr int kOffBar = (kOffDesc + kNS * kDescWords * 4 + 7) & ~7;   // full | scanned | free, kNS each
constexpr int kOffWsum = kOffBar + 3 * kNS * 8;               // [group][2][kSW]
constexpr int kOffQueue = kOffWsum + kSG * 2 * kSW * 4;       // [kQN] entries, then the head counter
constexpr int kOffLut = (kOffQueue + kQN * 4 + 16 + 15) & ~15;
// minimizer build: per row warp, kMinzW + 1 positions x 32 lanes x 8 bytes -- the hashes of the current block of W k-mers,
// turned into the block's suffix minima in place (two-level block minima; one lane per row, lanes in lock-step)
constexpr int kMinzW = 12;                      // longest window (in k-mers) this build takes
constexpr int kOffMinz = (kOffLut + 256 + 127) & ~127;
// ... and kStageU code words per lane: rows of up to kStageU units are encoded into them first, the slot goes back to
// the ring, and the (long) minimizer walk runs on the staged words -- the ring then only has to cover the front end
constexpr int kStageU = 13;
constexpr int kOffStage = kOffMinz + (MINZ ? kRW * (kMinzW + 1) * 256 : 0);   // + the all-ones position
constexpr int kFixedBytes = kOffStage + (MINZ ? kRW * kStageU * 128 : 0);
static_assert(kOffBar % 8 == 0 && kOffLut % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint; sleeping between the tests (nanosleep 64) costs more in wake-up latency than the
// polling costs in issue slots (1.53 -> 2.05 ms), so the loop polls.
#ifndef BNPK_MBAR_WAIT_BODY
#define BNPK_MBAR_WAIT_BODY                                                      \
    asm volatile(                                                                \
        "{\n"                                                                    \
        ".reg .pred p;\n"                                                        \
        "WAIT_%=:\n"                                                             \
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"            \
        "@p bra DONE_%=;\n"                                                      \
        "bra WAIT_%=;\n"                                                         \
        "DONE_%=:\n"                                                             \
        "}\n" ::"r"(bar), "r"(parity), "r"(20000u)                               \
        : "memory")
#endif
// one copy per waiting role, so that profiles tell the waits apart
__device__ __forceinline__ void mbar_wait_free(uint32_t bar, uint32_t parity) {
    BNPK_MBAR_WAIT_BODY;
}
__device__ __forceinline__ void mbar_wait_full_f(uint32_t bar, uint32_t parity) {
    BNPK_MBAR_WAIT_BODY;
}
__device__ __forceinline__ void mbar_wait_scanned(uint32_t bar, uint32_t parity) {
    BNPK_MBAR_WAIT_BODY;
}
__device__ __forceinline__ void mbar_wait_full_s(uint32_t bar, uint32_t parity) {
    BNPK_MBAR_WAIT_BODY;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void scan_bar(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kSW * 32) : "memory"); }
__device__ __forceinline__ uint4 lds128(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
// PRMT without the selector clean-up __byte_perm adds (all selectors used here are in range)
__device__ __forceinline__ uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}
// one count into the CTA-private table (32-bit shared address)
__device__ __forceinline__ void hist_inc(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
// ptxas never predicates ATOMS (it branches around it), so a masked count adds 0 or 1 instead
__device__ __forceinline__ void hist_add_val(uint32_t addr, uint32_t val) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory"); }

// Integer pipes of an SM sub-partition (tools/micro/pipe_bench.cu, B200): LOP3/SHF/PRMT/IADD3 (ALU pipe) and IMAD /
// IDP.4A (FMA pipe) each issue one warp instruction every two cycles, any mix of the two 0.65 per cycle; POPC one
// every 8 cycles, ffs (BREV + FLO) one every 16.  This path is all integer work: instruction count sets the time.

// bit 7 of every byte that equals '\n' (bit 7 of the pattern is clear, so the last term can use w itself)
__device__ __forceinline__ uint32_t newline_msb(uint32_t w) {
    uint32_t x;                                                     // (w ^ 0x0A..) & 0x7F.. as ONE LOP3
    asm("lop3.b32 %0, %1, 0x0A0A0A0A, 0x7F7F7F7F, 0x28;" : "=r"(x) : "r"(w));
    const uint32_t s = x + 0x7F7F7F7Fu;
    return ~(s | w) & 0x80808080u;
}
// exact '\n' flags of a 16-byte unit, bit i = byte i: the flag bytes are 0x80 or 0, one IDP.4A per word weighs
// them into place (4 instructions per word, two of them on the FMA pipe)
__device__ __forceinline__ uint32_t newline_mask16(const uint4 q) {
    uint32_t lo = __dp4a(newline_msb(q.x), 0x08040201u, 0u);
    lo = __dp4a(newline_msb(q.y), 0x80402010u, lo);                // 128 * (flags of bytes 0..7)
    uint32_t hi = __dp4a(newline_msb(q.z), 0x08040201u, 0u);
    hi = __dp4a(newline_msb(q.w), 0x80402010u, hi);                // 128 * (flags of bytes 8..15)
    return (lo >> 7) | (hi << 1);
}

// Conflict-free read of a lane's 64 bytes (LDS.128 j fetches unit (j + lane/2) & 3) -> exact 64-bit newline mask.
struct ScanLane {
    uint32_t off[4], sel_lo, sel_hi;
    __device__ __forceinline__ void init(int lane) {
        const uint32_t rot = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j) off[j] = 64u * (uint32_t)lane + 16u * (((uint32_t)j + rot) & 3u);
        // halfword h of the byte-order mask comes from load (h - rot) & 3; PRMT byte pair of load jj in
        // (A = m0|m1<<16, B = m2|m3<<16) is 0x10 + 0x22*jj
        sel_lo = (0x10u + 0x22u * ((0u - rot) & 3u)) | ((0x10u + 0x22u * ((1u - rot) & 3u)) << 8);
        sel_hi = (0x10u + 0x22u * ((2u - rot) & 3u)) | ((0x10u + 0x22u * ((3u - rot) & 3u)) << 8);
    }
    // p = base of the warp's 2 KiB piece
    __device__ __forceinline__ uint64_t mask64(const uint8_t *p) const {
        uint32_t m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = newline_mask16(lds128(p + off[j]));
        const uint32_t A = m[1] * 65536u + m[0], B = m[3] * 65536u + m[2];
        return ((uint64_t)prmt(A, B, sel_hi) << 32) | prmt(A, B, sel_lo);
    }
};

// write the positions of the set bits of m (tile-relative base `pos`) at list[li - wb ...] when inside the window
__device__ __forceinline__ void emit_positions(uint64_t m, uint32_t li, uint32_t pos, uint16_t *list, uint32_t wb) {
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    while (lo) {
        const int bit = __ffs((int)lo) - 1;
        lo &= lo - 1;
        if (li - wb < (uint32_t)kNlCap) list[li - wb] = (uint16_t)(pos + bit);
        ++li;
    }
    while (hi) {
        const int bit = __ffs((int)hi) - 1;
        hi &= hi - 1;
        if (li - wb < (uint32_t)kNlCap) list[li - wb] = (uint16_t)(pos + 32 + bit);
        ++li;
    }
}

// The scan warps' version for both pieces of a lane at once: four independent bit chains (low and high half of
// each 64-bit mask) advance together, so the latency of the bit searches (XU pipe) overlaps instead of adding up.
__device__ __forceinline__ void emit_positions2(uint64_t m0, uint32_t li0, uint32_t pos0, uint64_t m1, uint32_t li1, uint32_t pos1,
                                                uint16_t *list) {
    uint32_t a = (uint32_t)m0, b = (uint32_t)(m0 >> 32), c = (uint32_t)m1, d = (uint32_t)(m1 >> 32);
    uint32_t ia = li0, ib = li0 + (uint32_t)__popc(a), ic = li1, id = li1 + (uint32_t)__popc(c);
    while (a | b | c | d) {
        if (a) { const uint32_t l = a & (0u - a); a ^= l; if (ia < (uint32_t)kNlCap) list[ia] = (uint16_t)(pos0 + __popc(l - 1u)); ++ia; }
        if (b) { const uint32_t l = b & (0u - b); b ^= l; if (ib < (uint32_t)kNlCap) list[ib] = (uint16_t)(pos0 + 32u + __popc(l - 1u)); ++ib; }
        if (c) { const uint32_t l = c & (0u - c); c ^= l; if (ic < (uint32_t)kNlCap) list[ic] = (uint16_t)(pos1 + __popc(l - 1u)); ++ic; }
        if (d) { const uint32_t l = d & (0u - d); d ^= l; if (id < (uint32_t)kNlCap) list[id] = (uint16_t)(pos1 + 32u + __popc(l - 1u)); ++id; }
    }
}

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t lo, uint32_t hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
    return ((uint64_t)hi << 32) | lo;
}
// bins that are not a power of two: out of line, the callers' loops stay small
__device__ __noinline__ uint64_t mod_bins(uint64_t h, uint64_t n_bins) { return h % n_bins; }

// 16-byte unit -> 32 bits of 2-bit codes; `bad` != 0 iff a byte of the unit (WHOLE) or a byte selected by seq16
// (!WHOLE) is outside the alphabet (exact).
// ASCII alphabets: bits 1-2 of a letter are a Gray code of its index (A 00, C 01, G 11, T 10).  Per word: one LOP3
// isolates them, one IMAD packs the four fields into the top byte, one IMAD lines them up as PRMT selector nibbles,
// PRMT looks the expected lower-case letter up, LOP3 compares it with the case-folded input; per unit: three PRMT
// gather the packed bytes and (ACGT only) two ops turn Gray into binary for all sixteen bases at once.
template <int ENC, bool WHOLE>
__device__ __forceinline__ uint32_t encode_unit(const uint4 q, uint32_t seq16, const uint8_t *s_lut, uint32_t &bad) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    if constexpr (ENC == BNPK_ENC_ASCII_ACGT || ENC == BNPK_ENC_ASCII_ACTG || ENC == BNPK_ENC_CODES) {
        uint32_t dif[4], pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if constexpr (ENC == BNPK_ENC_CODES) {
                pk[j] = (w[j] & 0x03030303u) * 0x01041040u;
                dif[j] = w[j] & 0xFCFCFCFCu;
            } else {
                const uint32_t g2 = w[j] & 0x06060606u;
                pk[j] = g2 * 0x00820820u;                           // top byte = the four 2-bit fields
                const uint32_t sel = prmt(g2 * 0x110u, 0u, 0x4431u);   // nibbles = 2 * field: 0 a, 2 c, 4 t, 6 g
                dif[j] = prmt(0x00630061u, 0x00670074u, sel) ^ (w[j] | 0x20202020u);
            }
        }
        uint32_t codes = prmt(prmt(pk[0], pk[1], 0x0073), prmt(pk[2], pk[3], 0x0073), 0x5410);
        if constexpr (ENC == BNPK_ENC_ASCII_ACGT) codes ^= (codes >> 1) & 0x55555555u;
        if constexpr (WHOLE) {
            bad = dif[0] | dif[1] | dif[2] | dif[3];
        } else {
            // byte != 0 flags (bit 7 of every byte), weighed into a 16-bit mask like the newline flags
            uint32_t nz[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) nz[j] = (((dif[j] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dif[j]) & 0x80808080u;
            const uint32_t lo = __dp4a(nz[1], 0x80402010u, __dp4a(nz[0], 0x08040201u, 0u));
            const uint32_t hi = __dp4a(nz[3], 0x80402010u, __dp4a(nz[2], 0x08040201u, 0u));
            bad = ((lo >> 7) | (hi << 1)) & seq16;
        }
        return codes;
    } else {
        return encode_unit_seq<ENC>(w, WHOLE ? 0xFFFFu : seq16, s_lut, bad);
    }
}

// Checks of a tile that do not belong to one of its rows, by lanes 0..3 of a warp that owns the slot: entry
// structure at the newlines before the first row's (one_line_buffer.py:155-173, fastq_buffer.py:38-45), the
// chunk's first byte, the end of the tile's last complete entry (-> FileBuffer.size, one_line_buffer.py:67-69).
__device__ __forceinline__ void tile_head_checks(const TileArgs &a, const uint8_t *sp, const uint16_t *list, int64_t tile,
                                                 uint32_t count, uint64_t base, uint32_t ls, uint32_t want, int lane,
                                                 unsigned long long &complete) {
    const uint32_t pm = (1u << ls) - 1u;
    const size_t byte0 = (size_t)tile * kTileBytes;
    const uint32_t base_phase = (uint32_t)base & pm;
    const int64_t q0 = (int64_t)(base >> ls);
    const uint32_t jr0 = (want - base_phase) & pm;                   // first newline (rel) that precedes a field line
    if ((uint32_t)lane < jr0 && (uint32_t)lane < count) {
        const uint32_t p = list[lane];
        const uint32_t phase = (base_phase + (uint32_t)lane) & pm;
        const bool chk_h = phase == pm, chk_p = a.check_plus && phase == 1u;
        if ((chk_h || chk_p) && byte0 + p + 1 < a.n) {
            const uint32_t c = sp[p + 1];
            if (chk_h && c != a.header_char)
                atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], (long long)(q0 + ((base_phase + (uint32_t)lane + 1u) >> ls)));
            if (chk_p && c != '+')
                atomicMin((long long *)&a.status[BNPK_ST_BAD_PLUS_ENTRY], (long long)(q0 + ((base_phase + (uint32_t)lane) >> ls)));
        }
    }
    if (lane == 0) {
        if (count > 0) {                                              // last complete entry of the tile
            const uint32_t last = count - 1u;
            const uint32_t back = (base_phase + last - pm) & pm;
            if (last >= back && last - back < (uint32_t)kNlCap)
                complete = max(complete, (unsigned long long)(byte0 + list[last - back] + 1));
        }
        if (tile == 0 && a.n > 0 && sp[0] != a.header_char)
            atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], 0ll);
    }
}

// HIST: 0 = global int64 table, 1 = CTA-private u32 table in shared memory, 2 = global u32 scratch table
template <int ENC, int HIST>
__global__ void __launch_bounds__(kCta, 1) tile_ws_kernel(const TileArgs a) {
    constexpr bool SMEM_HIST = HIST == 1;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *s_fixed = smem_raw + (SMEM_HIST ? ((a.n_bins * 4 + 4 + 127) & ~(uint64_t)127) : 0);   // table + a spare word
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *s_slots = s_fixed + kOffSlots;
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_fixed + kOffList);
    volatile uint32_t *s_desc = reinterpret_cast<volatile uint32_t *>(s_fixed + kOffDesc);
    const uint32_t bar_full = smem_addr(s_fixed + kOffBar), bar_scanned = bar_full + 8 * kNS, bar_free = bar_full + 16 * kNS;
    volatile uint32_t *s_wsum = reinterpret_cast<volatile uint32_t *>(s_fixed + kOffWsum);
    volatile uint32_t *s_queue = reinterpret_cast<volatile uint32_t *>(s_fixed + kOffQueue);
    uint32_t *s_qhead = const_cast<uint32_t *>(s_queue) + kQN;
    uint8_t *s_lut = s_fixed + kOffLut;
    uint64_t *tile_state = a.ws + kWsHeaderWords;

    if (ENC == BNPK_ENC_LUT && tid < 256) s_lut[tid] = a.lut[tid];
    if (SMEM_HIST)
        for (uint32_t b = tid; b < a.n_bins; b += kCta) s_hist[b] = 0;
    if (tid < kQN + 4) s_queue[tid] = 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_scanned + 8 * s, kSW);
            mbar_init(bar_free + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t ls = (uint32_t)a.lpe_shift, pm = (1u << ls) - 1u;
    const uint32_t want = ((uint32_t)a.field_line - 1u) & pm;      // phase of the newline that precedes a field line
    const int32_t tile_end = (int32_t)a.tile_end;

    if (warp == kPWarp) {
        // ============================ P: tickets and bulk copies ============================================
        // Tiles are dealt round-robin over the CTAs (CTA c takes tiles c, c + G, c + 2G, ...): no ticket counter --
        // one global atomic per tile on a single address cost ~1.4 us of latency each and bounded the whole kernel.
        // Every CTA is resident (grid <= SM count), so the look-back never waits for a tile nobody has started.
        if (lane == 0) {
            int nend = 0;
            for (uint32_t seq = 0;; ++seq) {
                const uint32_t slot = seq % kNS, use = seq / kNS;
                const uint32_t tw0 = (BNPK_WS_DBG(a) & 16) ? (uint32_t)clock64() : 0u;
                if (use > 0) mbar_wait_free(bar_free + 8 * slot, (use - 1u) & 1u);
                if (BNPK_WS_DBG(a) & 16) {
                    atomicAdd((unsigned long long *)(a.ws + 9), (unsigned long long)((uint32_t)clock64() - tw0));
                    s_desc[slot * kDescWords + kDTIssue] = (uint32_t)clock64();
                }
                const int64_t t = a.tile_begin + (int64_t)blockIdx.x + (int64_t)seq * gridDim.x;
                if (t < (int64_t)tile_end) {
                    const size_t byte0 = (size_t)t * kTileBytes;
                    const uint32_t bytes = (uint32_t)min((size_t)kSlot, a.n - byte0) & ~15u;
                    s_desc[slot * kDescWords + kDTile] = (uint32_t)t;
                    if (bytes) {
                        mbar_expect_tx(bar_full + 8 * slot, bytes);
                        bulk_g2s(smem_addr(s_slots + slot * kSlot), a.chunk + byte0, bytes, bar_full + 8 * slot);
                    } else {
                        mbar_arrive(bar_full + 8 * slot);
                    }
                    // the tile this slot gets next: into L2 now, so that its copy later is an L2 hit (shorter ring
                    // latency: 1.535 -> 1.46 ms together with the debug knobs compiled out)
                    const int64_t tp = t + (int64_t)kNS * gridDim.x;
                    if (tp < (int64_t)tile_end) {
                        const size_t pb = (size_t)tp * kTileBytes;
                        const uint32_t pbytes = (uint32_t)min((size_t)kSlot, a.n - pb) & ~15u;
                        if (pbytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.chunk + pb), "r"(pbytes) : "memory");
                    }
                } else {                                              // one end marker per scan group
                    s_desc[slot * kDescWords + kDTile] = 0xFFFFFFFFu;
                    mbar_arrive(bar_full + 8 * slot);
                    if (++nend == kSG) break;
                }
            }
        }
    } else if (warp == kFWarp) {
        // ============================ F: line index of every tile of this CTA; deals the rows out ===========
        // A single warp per CTA, once per tile: its latency per tile bounds the whole ring, so the path from "the scan
        // warps are done" to "the row warps have their chunks" is kept to one shared-memory read and a few stores; the
        // sum over the other CTAs' counts is complete before that (loads issued one tile ahead, 32-bit REDUX).
        const int64_t G = (int64_t)gridDim.x;
        uint64_t incl = a.tile_begin > 0 ? a.ws[kWsCarry] : 0ull;    // lines in every tile before `prev + 1`
        int64_t prev = a.tile_begin - 1, tile = a.tile_begin + (int64_t)blockIdx.x;
        unsigned long long f_complete = 0;
        uint32_t pushes = 0;                                         // chunks queued so far (uniform)
        auto push = [&](uint32_t idx, uint32_t rec) {                // one lane per record
            volatile uint32_t *qe = s_queue + (idx & (kQN - 1));
            while (*qe != 0u) __nanosleep(64);                       // the entry's previous chunk has not been taken yet
            *qe = (((idx & 0x7FFFu) + 1u) << 16) | rec;
        };
        auto issue = [&](uint64_t *v, int64_t lo, int64_t hi) {      // counts of the tiles lo+1 .. hi-1
#pragma unroll
            for (int i = 0; i < kFK; ++i) {
                const int64_t idx = lo + 1 + lane + 32 * i;
                v[i] = (hi < (int64_t)tile_end && idx < hi) ? ld_relaxed(tile_state + idx) : kFlagAgg;
            }
        };
        uint64_t vC[kFK];
        issue(vC, prev, tile);
        for (uint32_t seq = 0; tile < (int64_t)tile_end; ++seq) {
            const uint32_t slot = seq % kNS, par = (seq / kNS) & 1u;
            uint64_t vN[kFK];
            issue(vN, tile, tile + G);                                // the next tile's, a tile period ahead
            uint32_t sum32 = 0;
#pragma unroll
            for (int i = 0; i < kFK; ++i) {
                uint64_t v = vC[i];
                const int64_t idx = prev + 1 + lane + 32 * i;
                while ((v >> 62) == 0) v = ld_relaxed(tile_state + idx);
                sum32 += (uint32_t)v;                                 // a tile has at most 16384 newlines
            }
            uint64_t base = incl + __reduce_add_sync(0xffffffffu, sum32);
            if (tile - prev - 1 > 32 * kFK) {                         // rare: more CTAs than the loads in flight cover
                uint64_t sum = 0;
                for (int64_t idx = prev + 1 + lane + 32 * kFK; idx < tile; idx += 32) {
                    uint64_t v = ld_relaxed(tile_state + idx);
                    while ((v >> 62) == 0) v = ld_relaxed(tile_state + idx);
                    sum += v & kValueMask;
                }
                base += warp_sum_u64(sum);
            }
            mbar_wait_scanned(bar_scanned + 8 * slot, par);
            const uint32_t count = s_desc[slot * kDescWords + kDCount];
            incl = base + count;
            const uint32_t jr0 = (want - (uint32_t)base) & pm;        // first newline (rel) that precedes a field line
            const uint32_t n_rows = count > jr0 ? ((count - 1u - jr0) >> ls) + 1u : 0u;
            const bool whole = count > (uint32_t)kNlCap;
            const uint32_t n_chunks = whole ? 1u : (n_rows + 31u) >> 5;
            if (lane == 0) {
                if (BNPK_WS_DBG(a) & 16) s_desc[slot * kDescWords + kDTPush] = (uint32_t)clock64();
                *reinterpret_cast<volatile uint64_t *>(s_desc + slot * kDescWords + kDBase) = base;
                s_desc[slot * kDescWords + kDRemain] = n_chunks;
            }
            __threadfence_block();
            __syncwarp();
            // the chunks go into the one queue every row warp takes from: whichever is free first gets the next
            for (uint32_t c = (uint32_t)lane; c < n_chunks; c += 32) push(pushes + c, (slot << 12) | (whole ? kChunkWhole : c));
            pushes += n_chunks;
            // off the critical path
            if (n_chunks == 0) {                                      // rare: no row starts in this tile; nobody else looks at it
                tile_head_checks(a, s_slots + slot * kSlot, s_list + slot * kNlCap, tile, count, base, ls, want, lane, f_complete);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_free + 8 * slot);
            }
            if (lane == 0) {
                if (tile == a.n_tiles_total - 1) a.status[BNPK_ST_N_LINES] = (int64_t)(base + count);
                if (tile == (int64_t)tile_end - 1) a.ws[kWsCarry] = base + count;
            }
#pragma unroll
            for (int i = 0; i < kFK; ++i) vC[i] = vN[i];
            prev = tile;
            tile += G;
        }
        if (lane < kRW) push(pushes + (uint32_t)lane, kChunkEnd);
        if (lane == 0 && f_complete)
            atomicMax((unsigned long long *)&a.status[BNPK_ST_N_COMPLETE_BYTES], f_complete);
    } else if (warp < kRWarp0) {
        // ============================ S: newline masks, sorted newline list, tile count ======================
        const int group = (warp - kSWarp0) / kSW, sw = (warp - kSWarp0) % kSW;
        ScanLane sl;
        sl.init(lane);
        for (uint32_t seq = (uint32_t)group;; seq += kSG) {
            const uint32_t slot = seq % kNS, par = (seq / kNS) & 1u;
            mbar_wait_full_s(bar_full + 8 * slot, par);
            const int32_t tile = (int32_t)s_desc[slot * kDescWords + kDTile];
            if (tile < 0) break;
            if ((BNPK_WS_DBG(a) & 16) && sw == 0 && lane == 0) s_desc[slot * kDescWords + kDTFull] = (uint32_t)clock64();
            uint8_t *sp = s_slots + slot * kSlot;
            const size_t byte0 = (size_t)tile * kTileBytes;
            const int staged = (int)min((size_t)kSlot, a.n - byte0);
            if (staged & 15) {                                        // the chunk's last bytes: not a multiple of 16
                const int t0 = staged & ~15;
                if (sw == 0 && lane < (staged & 15)) sp[t0 + lane] = a.chunk[byte0 + t0 + lane];
                scan_bar(group);
            }
            const uint8_t *pb = sp + 4096 * sw;
            uint64_t nl0 = sl.mask64(pb), nl1 = sl.mask64(pb + 2048);
            if (staged < kTileBytes) {                                // the chunk's last tile: bytes inside the chunk only
                const int lim0 = staged - (4096 * sw + 64 * lane), lim1 = lim0 - 2048;
                if (lim0 < 64) nl0 = lim0 <= 0 ? 0ull : (nl0 & (~0ull >> (64 - lim0)));
                if (lim1 < 64) nl1 = lim1 <= 0 ? 0ull : (nl1 & (~0ull >> (64 - lim1)));
            }
            const uint32_t cnt0 = (uint32_t)__popcll(nl0), cnt1 = (uint32_t)__popcll(nl1);
            uint32_t inc = cnt0 | (cnt1 << 16);                       // both pieces in one scan (a piece has <= 2048 newlines)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t t0 = tot & 0xFFFFu;
            volatile uint32_t *wsum = s_wsum + (group * 2 + ((seq / kSG) & 1u)) * kSW;
            if (lane == 0) wsum[sw] = t0 + (tot >> 16);
            if (sw == kSW - 1) {                                      // first newline of the halo: end of the crossing row
                const int valid = min(max(staged - kTileBytes - 16 * lane, 0), 16);
                const uint32_t mm = newline_mask16(lds128(sp + kTileBytes + 16 * lane)) & ((1u << valid) - 1u);
                const unsigned b = __ballot_sync(0xffffffffu, mm != 0);
                const int srcl = b ? __ffs(b) - 1 : 0;
                const uint32_t pos = (uint32_t)(kTileBytes + 16 * lane + __ffs(mm) - 1);
                const uint32_t first = __shfl_sync(0xffffffffu, pos, srcl);
                if (lane == 0) s_desc[slot * kDescWords + kDCross] = b ? first : kNoCross;
            }
            scan_bar(group);                                          // warp totals of this tile visible (double-buffered)
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kSW; ++w) {
                const uint32_t v = wsum[w];
                total += v;
                if (w < sw) before += v;
            }
            if (sw == 0 && lane == 0) {                               // the other CTAs wait for this: out before the list
                s_desc[slot * kDescWords + kDCount] = total;
                st_relaxed(tile_state + tile, kFlagAgg | (uint64_t)total);
            }
            uint16_t *list = s_list + slot * kNlCap;
            const uint32_t pos0 = 4096u * (uint32_t)sw + 64u * (uint32_t)lane;
            emit_positions2(nl0, before + (inc & 0xFFFFu) - cnt0, pos0, nl1, before + t0 + (inc >> 16) - cnt1, pos0 + 2048u, list);
            if ((BNPK_WS_DBG(a) & 16) && sw == 0 && lane == 0) s_desc[slot * kDescWords + kDTScanned] = (uint32_t)clock64();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_scanned + 8 * slot);
        }
    } else {
        // ============================ R: rows -> codes -> k-mers -> histogram =================================
        if constexpr (MINZ) {                                         // the all-ones position behind every window length
            for (int p = 0; p <= kMinzW; ++p)
                sts64(smem_addr(s_fixed + kOffMinz) + (uint32_t)((warp - kRWarp0) * ((kMinzW + 1) * 256) + 256 * p + 8 * lane), 0xFFFFFFFFu, 0xFFFFFFFFu);
            __syncwarp();
        }
        const bool cr = a.status[BNPK_ST_CR] != 0;
        const uint64_t hmask = (a.n_bins & (a.n_bins - 1)) == 0 ? a.n_bins - 1 : 0;
        const uint64_t kmask = (1ull << (2 * a.k)) - 1;
        [[maybe_unused]] const bool fast = hmask && hmask <= 0x3FFFFFFFull;
        const int dbg = BNPK_WS_DBG(a);                               // development knobs (BNPK_WS_DEBUG), 0 in production builds
        [[maybe_unused]] const uint32_t m32x4 = (dbg & 1) ? 0u : (uint32_t)(hmask & kmask) << 2;   // byte-offset mask into the table
        const uint32_t hist_sa = smem_addr(s_hist) + ((dbg & 1) ? 4u * (uint32_t)lane : 0u);
        uint32_t acc_bases = 0, acc_values = 0;                       // per lane: well inside 32 bits for any chunk
        unsigned long long last_start = 0, last_index = 0;             // 1 + start / entry of the last row this lane counted
        unsigned long long r_complete = 0;

        auto defer_row = [&](uint64_t start, uint64_t r) {
            const unsigned long long d = atomicAdd((unsigned long long *)(a.ws + kWsDeferred), 1ull);
            if (d < a.deferred_cap) {
                a.deferred[2 * d] = start;
                a.deferred[2 * d + 1] = r;
            } else {
                a.status[BNPK_ST_OVERFLOW] = 1;
            }
        };

        for (;;) {
            // ---- next chunk: whoever asks first
            uint32_t idx = 0;
            if (lane == 0) idx = atomicAdd(s_qhead, 1u);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            // One-word messages: the producer fences and stores a tagged record, the consumer polls the word, clears it and
            // fences (volatile accesses + fence.acq_rel.cta on both sides).  compute-sanitizer's racecheck follows
            // bar.sync / __syncwarp only: it reports these two lines and every access that is ordered THROUGH this
            // hand-off or through the mbarriers (list, descriptor, slot bytes) as hazards
            // (profiles/r02_compute_sanitizer.txt lists them).  Polling with shared-memory atomics instead slowed the kernel
            // from 1.53 to 2.0 ms (the polls queue behind the histogram atomics) and silences only these two lines.
            volatile uint32_t *qe = s_queue + (idx & (kQN - 1));
            uint32_t rec = *qe;
            while ((rec >> 16) != (idx & 0x7FFFu) + 1u) {
                __nanosleep(100);
                rec = *qe;
            }
            __syncwarp();
            if (lane == 0) *qe = 0u;                                  // taken
            __threadfence_block();
            const uint32_t chunk_id = rec & 0xFFFu;
            if (chunk_id == kChunkEnd) break;
            const uint32_t slot = (rec >> 12) & 0xFu;
            const uint32_t t_start = dbg ? (uint32_t)clock64() : 0u;
            if ((dbg & 16) && lane == 0) {
                const uint32_t ti = s_desc[slot * kDescWords + kDTIssue], tf = s_desc[slot * kDescWords + kDTFull],
                               tsc = s_desc[slot * kDescWords + kDTScanned], tp = s_desc[slot * kDescWords + kDTPush];
                if (chunk_id == 0u) {
                    atomicAdd((unsigned long long *)(a.ws + 4), (unsigned long long)(tf - ti));
                    atomicAdd((unsigned long long *)(a.ws + 5), (unsigned long long)(tsc - tf));
                    atomicAdd((unsigned long long *)(a.ws + 6), (unsigned long long)(tp - tsc));
                    atomicAdd((unsigned long long *)(a.ws + 10), 1ull);
                }
                atomicAdd((unsigned long long *)(a.ws + 7), (unsigned long long)(t_start - tp));
                atomicAdd((unsigned long long *)(a.ws + 11), 1ull);
            }
            const int32_t tile = (int32_t)s_desc[slot * kDescWords + kDTile];
            const uint32_t tile_nl = s_desc[slot * kDescWords + kDCount];
            const uint32_t crossM = s_desc[slot * kDescWords + kDCross];
            const int64_t line_base = (int64_t)*reinterpret_cast<volatile uint64_t *>(s_desc + slot * kDescWords + kDBase);
            const uint8_t *sp = s_slots + slot * kSlot;
            uint16_t *list = s_list + slot * kNlCap;
            const size_t byte0 = (size_t)tile * kTileBytes;
            const int staged = (int)min((size_t)kSlot, a.n - byte0);
            const uint32_t base_phase = (uint32_t)line_base & pm;
            const int64_t q0 = line_base >> ls;                       // entry index of the tile's first line
            const uint32_t jr0 = (want - base_phase) & pm;
            const int64_t r_first = q0 + ((base_phase + jr0 + 1u) >> ls);
            const int n_rows_tile = (tile_nl > jr0) ? (int)(((tile_nl - 1u - jr0) >> ls) + 1u) : 0;

            [[maybe_unused]] bool released = false;                   // minimizer build: the chunk gave its slot back early
            auto release_slot = [&]() {                               // the last chunk of a tile gives its slot back
                if (lane == 0) {
                    __threadfence_block();
                    const uint32_t old = atomicSub(const_cast<uint32_t *>(s_desc + slot * kDescWords + kDRemain), 1u);
                    if (old == 1u) {
                        __threadfence_block();
                        mbar_arrive(bar_free + 8 * slot);
                    }
                }
            };

            // 32 rows, one per lane: rows 32*c .. 32*c+31 of the tile; wb = first newline index held by the list
            auto do_chunk = [&](int c, uint32_t wb) {
                const int s = 32 * c + lane;                          // my row (tile-relative)
                bool act = s < n_rows_tile;
                const uint32_t j = jr0 + ((uint32_t)s << ls);         // the newline before my row
                int b0 = 0, e = 0;
                if (act) {
                    const uint16_t *lp = list + (j - wb);
                    b0 = (int)lp[0] + 1;
                    // entry structure (one_line_buffer.py:155-173, fastq_buffer.py:38-45) at the entry's other newlines
#pragma unroll
                    for (uint32_t d = 1; d <= 3; ++d) {
                        if (d <= pm && j + d < tile_nl) {
                            const uint32_t phase = (want + d) & pm;
                            const bool chk_h = phase == pm, chk_p = a.check_plus && phase == 1u;
                            if (chk_h || chk_p) {
                                const uint32_t p = lp[d];
                                if (byte0 + p + 1 < a.n) {
                                    const uint32_t ch = sp[p + 1];
                                    if (chk_h && ch != a.header_char)
                                        atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], (long long)(q0 + ((base_phase + j + d + 1u) >> ls)));
                                    if (chk_p && ch != '+')
                                        atomicMin((long long *)&a.status[BNPK_ST_BAD_PLUS_ENTRY], (long long)(q0 + ((base_phase + j + d) >> ls)));
                                }
                            }
                        }
                    }
                    if (j + 1u < tile_nl) {
                        e = lp[1];
                    } else if (crossM != kNoCross) {                  // the row ends in the halo
                        e = (int)crossM;
                    } else {                                          // not terminated inside the slot
                        if (byte0 + staged < a.n) defer_row(byte0 + b0, (uint64_t)(r_first + s));
                        act = false;                                  // (else: unterminated last line, not an entry)
                    }
                    if (act && cr && e > b0 && sp[e - 1] == '\r') e -= 1;
                    if (act && e - b0 > kRowMax) {
                        defer_row(byte0 + b0, (uint64_t)(r_first + s));
                        act = false;
                    }
                }
                const int L = act ? e - b0 : 0;
                const int npos = max(L - a.k + 1, 0);
                if (act) {
                    acc_bases += (uint32_t)L;
                    acc_values += MINZ ? (uint32_t)max(L - a.window + 1, 0) : (uint32_t)npos;
                    if (s == n_rows_tile - 1) {                       // the tile's last counted row (see uncount_kernel)
                        last_start = max(last_start, (unsigned long long)(byte0 + b0) + 1ull);
                        last_index = max(last_index, (unsigned long long)(r_first + s) + 1ull);
                    }
                }
                const int A0 = b0 >> 4;
                const int nu = L > 0 ? ((e - 1) >> 4) - A0 + 1 : 0;  // 16-byte units my row touches
                const uint32_t o = (uint32_t)b0 & 15u;
                const int R = __reduce_max_sync(0xffffffffu, nu);
                const uint8_t *up = sp + 16 * A0;

                uint32_t badacc = 0;                                  // != 0: some byte of my row is outside the alphabet
                // any unit of my row -> code word (0 beyond the row); only the row's own bytes are validated
                auto enc_masked = [&](int r) -> uint32_t {
                    if (r >= nu || (dbg & 4)) return 0u;
                    const uint4 qq = lds128(up + 16 * r);
                    const int lo = r == 0 ? (int)o : 0, hi = min(e - 16 * (A0 + r), 16);
                    const uint32_t seq16 = (0xFFFFu >> (16 - hi)) & (0xFFFFu << lo);
                    uint32_t bad;
                    const uint32_t codes = encode_unit<ENC, false>(qq, seq16, s_lut, bad);
                    badacc |= bad;
                    return codes;
                };
                const uint32_t sh = 2u * o;                           // my row's stream starts at bit 2*o of its first word

                auto report_bad = [&]() {                             // rare: exact position of the first bad byte
                    for (int pp = b0; pp < e; ++pp) {
                        const uint32_t cc = sp[pp];
                        bool okb;
                        if (ENC == BNPK_ENC_CODES) okb = cc < 4;
                        else if (ENC == BNPK_ENC_LUT) okb = s_lut[cc] < 4;
                        else { const uint32_t uu = cc | 0x20u; okb = (uu == 'a' || uu == 'c' || uu == 'g' || uu == 't'); }
                        if (!okb) {
                            atomicMin((long long *)&a.status[BNPK_ST_BAD_BASE], (long long)(((r_first + s) << 32) | (int64_t)(pp - b0)));
                            break;
                        }
                    }
                };
                if constexpr (MINZ) {
                    // ---- minimizers (sequence/minimizers.py:8-17,50-54): the minimum hash of every window of W consecutive
                    // k-mers, counted.  Two-level block minima: the hashes go in blocks of W; the window that ends at position
                    // r of a block is (suffix r+1.. of the previous block) + (prefix ..r of this block).  The prefix minimum is
                    // a register; a block's hashes are stored as they come and turned into suffix minima in place when the
                    // block is complete.  The position is the same in every lane (all rows start at k-mer 0), so every branch
                    // below is warp-uniform; lanes past their row's end count into nothing.
                    const int W = a.window - a.k + 1;
                    // 32-bit shared addresses: position p of my lane is mb0 + 256 p; position W holds all ones for good (the
                    // "suffix" the last position of a block asks for: none)
                    const uint32_t mb0 = smem_addr(s_fixed + kOffMinz) + (uint32_t)((warp - kRWarp0) * ((kMinzW + 1) * 256) + 8 * lane);
                    const uint32_t mb_end = mb0 + 256u * (uint32_t)W;
                    const int maxnpos = __reduce_max_sync(0xffffffffu, npos);
                    const int nblk = (maxnpos + 15) >> 4;
                    const uint32_t nout = (uint32_t)max(npos - W + 1, 0);   // windows of my row
                    // Rows of up to kStageU units (all rows of the chunk: a warp-uniform choice): every unit is encoded and
                    // validated now, the words wait in this lane's staging column, and the slot goes back to the ring before
                    // the walk below, which is many times longer than everything before it.  Longer rows (and the whole-tile
                    // walk, which owns the slot for all its chunks) read the slot as they go.
                    const int n_words = max(R, nblk + 3);
                    const bool stage = chunk_id != kChunkWhole && n_words <= kStageU;
                    uint32_t *stg = reinterpret_cast<uint32_t *>(s_fixed + kOffStage) + (warp - kRWarp0) * (kStageU * 32) + lane;
                    if (stage) {
                        for (int u = 0; u < n_words; ++u) stg[u * 32] = enc_masked(u);
                        if (badacc) {
                            report_bad();
                            badacc = 0;
                        }
                        __syncwarp();
                        release_slot();
                        released = true;
                    }
                    auto word = [&](int u) -> uint32_t { return stage ? stg[u * 32] : enc_masked(u); };
                    uint32_t w0 = word(0), w1 = word(1), w2 = word(2);
                    uint32_t mb_r = mb0, ic = (uint32_t)(1 - W);      // ic: index of the window that ends at this step
                    uint64_t pre = ~0ull, sfx_next = ~0ull;           // sfx_next: the previous block's suffix minimum this step
                                                                      // needs, loaded a step ahead
                    for (int b = 0; b < nblk; ++b) {
                        const uint32_t w3 = word(b + 3);
                        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
                        const int tn = min(16, maxnpos - 16 * b);
#pragma unroll 2
                        for (int t = 0; t < tn; ++t) {                // compact on purpose: unrolled 16 times the loop does not
                            const uint32_t lo32 = __funnelshift_r(a0, a1, 2 * t) & (uint32_t)kmask;   // fit the instruction cache
                            const uint32_t hi32 = __funnelshift_r(a1, a2, 2 * t) & (uint32_t)(kmask >> 32);
                            const uint64_t h = ((uint64_t)hi32 << 32) | lo32;
                            sts64(mb_r, lo32, hi32);
                            pre = h < pre ? h : pre;
                            const uint64_t m = sfx_next < pre ? sfx_next : pre;   // before the first full window: not counted
                            const uint64_t bb = hmask ? (m & hmask) : mod_bins(m, a.n_bins);
                            const bool mine = ic < nout;
                            if constexpr (SMEM_HIST) hist_add_val(hist_sa + 4u * (uint32_t)bb, mine ? 1u : 0u);
                            else if (mine) {
                                if constexpr (HIST == 2) atomicAdd(a.hist32 + bb, 1u);
                                else atomicAdd(a.hist + bb, 1ull);
                            }
                            ++ic;
                            mb_r += 256u;
                            if (mb_r == mb_end) {                     // block complete: its suffix minima, in place; every
                                uint64_t sfx = ~0ull;                 // load is issued one element ahead of its use
                                uint32_t q = mb_end - 256u;
                                uint64_t nx = lds64(q);
                                for (; q > mb0; q -= 256u) {
                                    const uint64_t v = nx;
                                    if (q > mb0 + 256u) nx = lds64(q - 256u);
                                    sfx = v < sfx ? v : sfx;
                                    sts64(q, (uint32_t)sfx, (uint32_t)(sfx >> 32));
                                }
                                sfx_next = sfx;                       // suffix 1.. of the block: what position 0 needs (W = 1: none)
                                mb_r = mb0;
                                pre = ~0ull;
                            } else {
                                sfx_next = lds64(mb_r + 256u);        // still the previous block's (garbage before there is one:
                            }                                         // those windows are not counted)
                        }
                        w0 = w1;
                        w1 = w2;
                        w2 = w3;
                    }
                    if (!stage)
                        for (int u = nblk + 3; u < R; ++u) enc_masked(u);   // units no k-mer reaches are still validated
                } else if (fast) {
                    // A_b = the 16 bases from row position 16b on = funnel(w_b, w_b+1, 2o).  K-mer t of block b is the
                    // field at bit 2t of (A_b, A_b+1); shifted two bits less, (window & mask) is the table's byte offset.
                    uint32_t w1 = enc_masked(1);
                    uint32_t A_cur = __funnelshift_r(enc_masked(0), w1, sh);
                    int b = 0;
                    {
                        // ---- steady state: blocks that are full in every row of the chunk, units that are interior in
                        // every row.  One straight-line body per block: no votes, no branches, the next unit's load in flight.
                        const int bfull = __reduce_min_sync(0xffffffffu, nu > 0 ? npos >> 4 : 0x7FFFFFFF);
                        const int minnu = __reduce_min_sync(0xffffffffu, nu > 0 ? nu : 0x7FFFFFFF);
                        const int bs = (dbg & 6) ? 0 : min(min(bfull, minnu - 3), R);
                        // rows that are not there count into a spare word behind the table
                        const uint32_t mk = nu > 0 ? m32x4 : 0u, dm = nu > 0 ? 0u : (uint32_t)(a.n_bins * 4);
                        if (bs > 0) {
                            uint4 qn = lds128(up + 32);
                            for (; b < bs; ++b) {
                                const uint4 qq = qn;
                                qn = lds128(up + 16 * (b + 3));       // inside my row: b + 3 <= minnu - 1
                                uint32_t bad;
                                const uint32_t w2 = encode_unit<ENC, true>(qq, 0xFFFFu, s_lut, bad);
                                badacc |= bad;
                                const uint32_t A_nxt = __funnelshift_r(w1, w2, sh);
                                if constexpr (SMEM_HIST) {
#pragma unroll
                                    for (int t = 0; t < 16; ++t) {
                                        const uint32_t win = t == 0 ? A_cur << 2 : __funnelshift_r(A_cur, A_nxt, 2 * t - 2);
                                        hist_inc(hist_sa + ((win & mk) | dm));
                                    }
                                } else if (nu > 0) {                  // global table: one RED per k-mer, no per-k-mer test
#pragma unroll
                                    for (int t = 0; t < 16; ++t) {
                                        const uint32_t v = (t == 0 ? A_cur << 2 : __funnelshift_r(A_cur, A_nxt, 2 * t - 2)) & m32x4;
                                        if constexpr (HIST == 2) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.hist32) + v), 1u);
                                        else atomicAdd(reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(a.hist) + 2 * (size_t)v), 1ull);
                                    }
                                }
                                A_cur = A_nxt;
                                w1 = w2;
                            }
                        }
                    }
                    // ---- the rest: masked units, blocks that are not full everywhere
                    for (; b < R; ++b) {
                        const uint32_t w2 = enc_masked(b + 2);
                        const uint32_t A_nxt = __funnelshift_r(w1, w2, sh);
                        const int left = npos - 16 * b;               // k-mers that start in this block
                        if (__any_sync(0xffffffffu, left > 0) && !(dbg & 2)) {
                            if constexpr (SMEM_HIST) {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {      // two half blocks of eight
                                    if (__all_sync(0xffffffffu, left >= 8 * hb + 8)) {   // every lane has all eight
#pragma unroll
                                        for (int t = 8 * hb; t < 8 * hb + 8; ++t)
                                            hist_inc(hist_sa + ((t == 0 ? A_cur << 2 : __funnelshift_r(A_cur, A_nxt, 2 * t - 2)) & m32x4));
                                    } else if (__any_sync(0xffffffffu, left > 8 * hb)) {
#pragma unroll
                                        for (int t = 8 * hb; t < 8 * hb + 8; ++t)
                                            hist_add_val(hist_sa + ((t == 0 ? A_cur << 2 : __funnelshift_r(A_cur, A_nxt, 2 * t - 2)) & m32x4), (uint32_t)(t - left) >> 31);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int t = 0; t < 16; ++t) {
                                    const uint32_t v = (t == 0 ? A_cur << 2 : __funnelshift_r(A_cur, A_nxt, 2 * t - 2)) & m32x4;
                                    if (t < left) {
                                        if constexpr (HIST == 2) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.hist32) + v), 1u);
                                        else atomicAdd(reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(a.hist) + 2 * (size_t)v), 1ull);
                                    }
                                }
                            }
                        }
                        A_cur = A_nxt;
                        w1 = w2;
                    }
                } else {
                    // 64-bit hashes (any number of bins): block b reads words b .. b+3
                    uint32_t w0 = enc_masked(0), w1 = enc_masked(1), w2 = enc_masked(2);
                    for (int b = 0; b < R; ++b) {
                        const uint32_t w3 = enc_masked(b + 3);
                        const int left = npos - 16 * b;
                        if (left > 0) {
                            const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh),
                                           a2 = __funnelshift_r(w2, w3, sh);
#pragma unroll 4
                            for (int t = 0; t < 16; ++t) {
                                if (t < left) {
                                    const uint32_t lo32 = __funnelshift_r(a0, a1, 2 * t), hi32 = __funnelshift_r(a1, a2, 2 * t);
                                    const uint64_t h = (((uint64_t)hi32 << 32) | lo32) & kmask;
                                    const uint64_t bb = hmask ? (h & hmask) : (h % a.n_bins);
                                    if constexpr (HIST == 2) atomicAdd(a.hist32 + bb, 1u);
                                    else if constexpr (HIST == 1) atomicAdd(s_hist + (uint32_t)bb, 1u);
                                    else atomicAdd(a.hist + bb, 1ull);
                                }
                            }
                        }
                        w0 = w1;
                        w1 = w2;
                        w2 = w3;
                    }
                }
                if (badacc) report_bad();
            };

            if (chunk_id == 0u || chunk_id == kChunkWhole)           // the tile's own checks, once per tile
                tile_head_checks(a, sp, list, tile, tile_nl, (uint64_t)line_base, ls, want, lane, r_complete);
            if constexpr (MINZ) {
                // One call site for the chunk walk (the minimizer loop is large: two copies cost instruction-cache misses; the
                // k-mer build keeps its two call sites below -- with one it ran at 1.72 ms instead of 1.46).  Usually
                // one chunk of the list the scan warps left; rare: more newlines than the list holds -- then this warp
                // walks the whole tile in windows of the list, which it rebuilds itself from the bytes.
                const bool whole = chunk_id == kChunkWhole;
                const int n_chunks_t = (n_rows_tile + 31) >> 5, cpw = kWinRows >> 5;   // chunks per window
                const int wk_end = whole ? (n_chunks_t + cpw - 1) / cpw : 1;
                for (int wk = 0; wk < wk_end; ++wk) {
                    const uint32_t wb = whole ? (uint32_t)(wk * kWinRows) << ls : 0u;
                    if (whole) {
                        ScanLane sl;
                        sl.init(lane);
                        __syncwarp();
                        uint32_t running = 0;
                        for (int blk = 0; blk < kTileBytes / 2048; ++blk) {
                            uint64_t nl = sl.mask64(sp + 2048 * blk);
                            const int lim = min(staged, kTileBytes) - (2048 * blk + 64 * lane);
                            if (lim < 64) nl = lim <= 0 ? 0ull : (nl & (~0ull >> (64 - lim)));
                            const uint32_t cnt = (uint32_t)__popcll(nl);
                            uint32_t inc = cnt;
#pragma unroll
                            for (int oo = 1; oo < 32; oo <<= 1) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, oo);
                                if (lane >= oo) inc += t;
                            }
                            emit_positions(nl, running + inc - cnt, 2048u * (uint32_t)blk + 64u * (uint32_t)lane, list, wb);
                            running += __shfl_sync(0xffffffffu, inc, 31);
                        }
                        __syncwarp();
                        if (lane == 0 && wk > 0) {                    // last complete entry of the tile, if this window holds it
                            const uint32_t last = tile_nl - 1u;
                            const uint32_t back = (base_phase + last - pm) & pm;
                            if (last >= back && last - back >= wb && last - back - wb < (uint32_t)kNlCap)
                                r_complete = max(r_complete, (unsigned long long)(byte0 + list[last - back - wb] + 1));
                        }
                    }
                    const int c0 = whole ? wk * cpw : (int)chunk_id, c1 = whole ? min(n_chunks_t, (wk + 1) * cpw) : (int)chunk_id + 1;
                    for (int c = c0; c < c1; ++c) do_chunk(c, wb);
                }
            } else if (dbg & 8) {
            } else if (chunk_id != kChunkWhole) {
                do_chunk((int)chunk_id, 0u);
            } else {
                // rare: more newlines than the list holds.  This warp walks the tile in windows of the list, which it
                // rebuilds itself from the bytes (window 0 is what the scan warps left).
                ScanLane sl;
                sl.init(lane);
                const int n_chunks = (n_rows_tile + 31) >> 5, cpw = kWinRows >> 5;   // chunks per window
                for (int wk = 0; wk * cpw < n_chunks; ++wk) {
                    const uint32_t wb = (uint32_t)(wk * kWinRows) << ls;
                    if (wk > 0) {
                        __syncwarp();
                        uint32_t running = 0;
                        for (int blk = 0; blk < kTileBytes / 2048; ++blk) {
                            uint64_t nl = sl.mask64(sp + 2048 * blk);
                            const int lim = min(staged, kTileBytes) - (2048 * blk + 64 * lane);
                            if (lim < 64) nl = lim <= 0 ? 0ull : (nl & (~0ull >> (64 - lim)));
                            const uint32_t cnt = (uint32_t)__popcll(nl);
                            uint32_t inc = cnt;
#pragma unroll
                            for (int oo = 1; oo < 32; oo <<= 1) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, oo);
                                if (lane >= oo) inc += t;
                            }
                            emit_positions(nl, running + inc - cnt, 2048u * (uint32_t)blk + 64u * (uint32_t)lane, list, wb);
                            running += __shfl_sync(0xffffffffu, inc, 31);
                        }
                        __syncwarp();
                    }
                    if (lane == 0 && wk > 0) {                        // last complete entry of the tile, if this window holds it
                        const uint32_t last = tile_nl - 1u;
                        const uint32_t back = (base_phase + last - pm) & pm;
                        if (last >= back && last - back >= wb && last - back - wb < (uint32_t)kNlCap)
                            r_complete = max(r_complete, (unsigned long long)(byte0 + list[last - back - wb] + 1));
                    }
                    for (int c = wk * cpw; c < min(n_chunks, (wk + 1) * cpw); ++c) do_chunk(c, wb);
                }
            }
            if ((dbg & 16) && lane == 0) atomicAdd((unsigned long long *)(a.ws + 8), (unsigned long long)((uint32_t)clock64() - t_start));
            __syncwarp();
            if (!released) release_slot();
        }
        const uint64_t sum_bases = warp_sum_u64(acc_bases), sum_values = warp_sum_u64(acc_values);
#pragma unroll
        for (int oo = 16; oo; oo >>= 1) {
            last_start = max(last_start, __shfl_xor_sync(0xffffffffu, last_start, oo));
            last_index = max(last_index, __shfl_xor_sync(0xffffffffu, last_index, oo));
        }
        if (lane == 0) {
            if (sum_bases) atomicAdd((unsigned long long *)&a.status[BNPK_ST_N_BASES], sum_bases);
            if (sum_values) atomicAdd((unsigned long long *)&a.status[BNPK_ST_N_VALUES], sum_values);
            if (last_start) atomicMax((unsigned long long *)&a.status[BNPK_ST_LAST_ROW_START], last_start);
            if (last_index) atomicMax((unsigned long long *)&a.status[BNPK_ST_LAST_ROW_INDEX], last_index);
            if (r_complete) atomicMax((unsigned long long *)&a.status[BNPK_ST_N_COMPLETE_BYTES], r_complete);
        }
    }

    // ---- flush ---------------------------------------------------------------------------------------------
    if (SMEM_HIST) {
        __syncthreads();
        for (uint32_t b = tid; b < a.n_bins; b += kCta) {
            const uint32_t c = s_hist[b];
            if (c) atomicAdd(a.hist + b, (unsigned long long)c);
        }
    }
}

template <int ENC, int HIST>
static int launch_t(const TileArgs &a, cudaStream_t st) {
    auto kern = tile_ws_kernel<ENC, HIST>;
    const size_t smem = (size_t)kFixedBytes + (HIST == 1 ? ((a.n_bins * 4 + 4 + 127) & ~(uint64_t)127) : 0);
    BNPK_DYN_SMEM(kern, kFixedBytes + kMaxBins * 4 + 128);
    const int64_t n_tiles = a.tile_end - a.tile_begin;
    if (n_tiles <= 0) return 0;
    const int64_t grid = std::min<int64_t>(n_tiles, (int64_t)sm_count());
    TileArgs b = a;
    static const int dbg = [] { const char *e = std::getenv("BNPK_WS_DEBUG"); return e ? atoi(e) : 0; }();
    b.start_offset = dbg;
    profile_before(st);
    kern<<<(unsigned)grid, kCta, smem, st>>>(b);
    profile_after(st);
    BNPK_LAUNCHED("tile_ws_kernel");
    return 0;
}

template <int ENC>
static int launch_enc(const TileArgs &a, bool smem_hist, cudaStream_t st) {
    if constexpr (MINZ) return launch_t<ENC, 1>(a, st);           // CTA-private tables only (see wsm_count_eligible)
    else {
        if (smem_hist) return launch_t<ENC, 1>(a, st);
        return a.hist32 ? launch_t<ENC, 2>(a, st) : launch_t<ENC, 0>(a, st);
    }
}

}  // namespace BNPK_WS_NAMESPACE

int BNPK_WS_LAUNCH(const TileArgs &a, int enc_mode, bool smem_hist, cudaStream_t st) {
    switch (enc_mode) {
        case BNPK_ENC_ASCII_ACGT: return BNPK_WS_NAMESPACE::launch_enc<BNPK_ENC_ASCII_ACGT>(a, smem_hist, st);
        case BNPK_ENC_ASCII_ACTG: return BNPK_WS_NAMESPACE::launch_enc<BNPK_ENC_ASCII_ACTG>(a, smem_hist, st);
        case BNPK_ENC_CODES: return BNPK_WS_NAMESPACE::launch_enc<BNPK_ENC_CODES>(a, smem_hist, st);
        case BNPK_ENC_LUT: return BNPK_WS_NAMESPACE::launch_enc<BNPK_ENC_LUT>(a, smem_hist, st);
    }
    return set_err(BNPK_E_BADARG, "bad enc_mode");
}

}  // namespace bnpk
