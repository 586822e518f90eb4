// tile_ws_kernel.cu -- fused count (K6), warp-specialised: the dominant kernel of the hot path.
//
// One persistent CTA per SM.  The chunk streams through a ring of kNS shared-memory slots (16 KiB tile + 512 B of
// halo so the one row that crosses the tile end can finish in the slot) filled by cp.async.bulk (TMA).  Four kinds
// of warps work on different tiles of the ring at the same time and hand slots on through mbarriers only -- there is
// no CTA- or group-wide barrier on the path of a tile, and no warp executes another role's code:
//
//   P  (1 lane)   takes tile tickets in order, waits for a slot to be free, starts its bulk copy            -> full[slot]
//   S  (4 warps)  exact newline masks of the tile (128 B per lane), prefix over the four warps (the only named
//                 barrier: the four scan warps), sorted newline list of the tile, the tile's newline count to the
//                 workspace (one relaxed store), first newline of the halo                                  -> scanned[slot]
//   F  (1 warp)   line index of the tile's first byte: this CTA's previous tile + the counts every CTA published
//                 for the tiles in between (loads issued one tile ahead); entry structure of the newlines before
//                 the first row, last complete entry, first byte of the chunk                               -> ready[slot]
//   R  (16 warps, 4 teams of 4; team t takes every 4th tile of the ring) per chunk of 8 rows: one lane per newline
//                 validates '@' / '+', four lanes per row read the row's 16-byte units from the slot, encode + validate
//                 them, pass the 2-bit code words round with shuffles; every k-mer is SHF + LOP3 + ATOMS   -> free[slot]
//
// The warps with the most urgent work have the highest warp ids (the SM arbiter favours them): P, F, S, then R.
// Replaces io/one_line_buffer.py:44-71,139-182 + encodings/alphabet_encoding.py:34-46 + sequence/kmers.py:105-126 +
// sequence/count_encoded.py:173-177 in one pass over the chunk bytes.
#include "tile_common.cuh"

namespace bnpk {
namespace ws {

constexpr int kNS = 8;                          // ring slots
constexpr int kHalo = 512;
constexpr int kSlot = kTileBytes + kHalo;
constexpr int kNlCap = 1024;                    // newline positions of one tile kept in shared memory
constexpr int kWinStep = 960;                   // tiles with more newlines are walked in windows of the list
constexpr int kRowMax = 1024;                   // longer rows go to the deferred (one warp per segment) pass
constexpr int kMaxBins = 16384;
constexpr uint32_t kNoCross = 0xFFFFFFFFu;
constexpr int kSW = 4;                          // scan warps: 4 KiB of the tile each, 128 B per lane
constexpr int kTeams = 4, kTeamWarps = 4;       // row warps
constexpr int kRW = kTeams * kTeamWarps;
constexpr int kFWarp = kRW + kSW, kPWarp = kFWarp + 1;
constexpr int kWarps = kPWarp + 1;
constexpr int kCta = kWarps * 32;
constexpr int kFK = 6;                          // look-back loads per lane kept in flight (192 tiles)
static_assert(kSW * 4096 == kTileBytes, "scan geometry");
static_assert((kNS & (kNS - 1)) == 0, "ring size");

// per-slot descriptor (32-bit words)
constexpr int kDTile = 0;                       // P: tile index, -1 = end of the launch
constexpr int kDCount = 1;                      // S: newlines in the tile proper
constexpr int kDCross = 2;                      // S: first newline of the halo (slot-relative) or kNoCross
constexpr int kDBase = 4;                       // F: int64 line index of the tile's first byte
constexpr int kDescWords = 8;
// shared memory after the histogram (bytes)
constexpr int kOffSlots = 0;
constexpr int kOffList = kOffSlots + kNS * kSlot;
constexpr int kOffDesc = kOffList + kNS * kNlCap * 2;
constexpr int kOffBar = kOffDesc + kNS * kDescWords * 4;      // full | scanned | ready | free, kNS each
constexpr int kOffWsum = kOffBar + 4 * kNS * 8;
constexpr int kOffLut = kOffWsum + 2 * kSW * 4;
constexpr int kFixedBytes = kOffLut + 256;

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void scan_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kSW * 32) : "memory"); }
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
// IDP (FMA pipe) each issue one warp instruction every two cycles.  This path is all integer work: instruction
// count, and how it splits over the two pipes, sets the kernel time -- not bytes.

// bit 7 of every byte that equals '\n' (bit 7 of the pattern is clear, so the last term can use w itself)
__device__ __forceinline__ uint32_t newline_msb(uint32_t w) {
    uint32_t x;                                                     // (w ^ 0x0A..) & 0x7F.. as ONE LOP3
    asm("lop3.b32 %0, %1, 0x0A0A0A0A, 0x7F7F7F7F, 0x28;" : "=r"(x) : "r"(w));
    const uint32_t s = x + 0x7F7F7F7Fu;
    return ~(s | w) & 0x80808080u;
}
// exact '\n' flags of a 16-byte unit, bit i = byte i
__device__ __forceinline__ uint32_t newline_mask16(const uint4 q) {
#ifdef BNPK_WS_DP4A
    // the flag bytes are 0x80 or 0: one IDP.4A per word weighs them into place (FMA pipe)
    uint32_t lo = __dp4a(newline_msb(q.x), 0x08040201u, 0u);
    lo = __dp4a(newline_msb(q.y), 0x80402010u, lo);                // 128 * (flags of bytes 0..7)
    uint32_t hi = __dp4a(newline_msb(q.z), 0x08040201u, 0u);
    hi = __dp4a(newline_msb(q.w), 0x80402010u, hi);                // 128 * (flags of bytes 8..15)
    return (lo >> 7) | (hi << 1);
#else
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t acc = 0;
#pragma unroll
    for (int j = 3; j >= 0; --j) acc = __funnelshift_l(newline_msb(w[j]) * 0x00204081u, acc, 4);
    return acc & 0xFFFFu;
#endif
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

// 16-byte unit -> 32 bits of 2-bit codes; `bad` != 0 iff a byte selected by seq16 is outside the alphabet (exact).
// ASCII alphabets: bits 1-2 of a letter are a Gray code of its index (A 00, C 01, G 11, T 10).  Per word: one LOP3
// isolates them, one IMAD packs the four fields into the top byte, one IMAD lines them up as PRMT selector nibbles,
// PRMT looks the expected lower-case letter up, LOP3 compares it with the case-folded input; per unit: three PRMT
// gather the packed bytes and (ACGT only) two ops turn Gray into binary for all sixteen bases at once.
template <int ENC>
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
        if (seq16 == 0xFFFFu) {
            bad = dif[0] | dif[1] | dif[2] | dif[3];
        } else {
            uint32_t acc = 0;                                       // bit i = byte i of the unit differs
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const uint32_t nz = (((dif[j] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dif[j]) & 0x80808080u;  // byte != 0
                acc = __funnelshift_l(nz * 0x00204081u, acc, 4);
            }
            bad = acc & seq16;
        }
        return codes;
    } else {
        return encode_unit_seq<ENC>(w, seq16, s_lut, bad);
    }
}

// HIST: 0 = global int64 table, 1 = CTA-private u32 table in shared memory, 2 = global u32 scratch table
template <int ENC, int HIST>
__global__ void __launch_bounds__(kCta, 1) tile_ws_kernel(const TileArgs a) {
    constexpr bool SMEM_HIST = HIST == 1;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *s_fixed = smem_raw + (SMEM_HIST ? ((a.n_bins * 4 + 127) & ~(uint64_t)127) : 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *s_slots = s_fixed + kOffSlots;
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_fixed + kOffList);
    volatile uint32_t *s_desc = reinterpret_cast<volatile uint32_t *>(s_fixed + kOffDesc);
    const uint32_t bar_full = smem_addr(s_fixed + kOffBar), bar_scanned = bar_full + 8 * kNS,
                   bar_ready = bar_full + 16 * kNS, bar_free = bar_full + 24 * kNS;
    volatile uint32_t *s_wsum = reinterpret_cast<volatile uint32_t *>(s_fixed + kOffWsum);
    uint8_t *s_lut = s_fixed + kOffLut;
    uint64_t *tile_state = a.ws + kWsHeaderWords;

    if (ENC == BNPK_ENC_LUT && tid < 256) s_lut[tid] = a.lut[tid];
    if (SMEM_HIST)
        for (uint32_t b = tid; b < a.n_bins; b += kCta) s_hist[b] = 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_scanned + 8 * s, kSW);
            mbar_init(bar_ready + 8 * s, 1);
            mbar_init(bar_free + 8 * s, kTeamWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t ls = (uint32_t)a.lpe_shift, pm = (1u << ls) - 1u;
    const uint32_t want = ((uint32_t)a.field_line - 1u) & pm;      // phase of the newline that precedes a field line
    const int32_t tile_end = (int32_t)a.tile_end;

    if (warp == kPWarp) {
        // ============================ P: tickets and bulk copies ============================================
        if (lane == 0) {
            auto take_ticket = [&]() -> int32_t {
                const unsigned long long t = (unsigned long long)a.tile_begin + atomicAdd((unsigned long long *)(a.ws + kWsTicket), 1ull);
                return (int32_t)min(t, (unsigned long long)0x7FFFFFFF);
            };
            int nend = 0;
            int32_t t = take_ticket();
            for (uint32_t seq = 0;; ++seq) {
                const uint32_t slot = seq & (kNS - 1), use = seq / kNS;
                if (use > 0) mbar_wait(bar_free + 8 * slot, (use - 1u) & 1u);
                if (t < tile_end) {
                    const int32_t t_next = take_ticket();            // in flight while this tile's copy is issued
                    const size_t byte0 = (size_t)t * kTileBytes;
                    const uint32_t bytes = (uint32_t)min((size_t)kSlot, a.n - byte0) & ~15u;
                    s_desc[slot * kDescWords + kDTile] = (uint32_t)t;
                    if (bytes) {
                        mbar_expect_tx(bar_full + 8 * slot, bytes);
                        bulk_g2s(smem_addr(s_slots + slot * kSlot), a.chunk + byte0, bytes, bar_full + 8 * slot);
                    } else {
                        mbar_arrive(bar_full + 8 * slot);
                    }
                    t = t_next;
                } else {                                              // one end marker per row team
                    s_desc[slot * kDescWords + kDTile] = 0xFFFFFFFFu;
                    mbar_arrive(bar_full + 8 * slot);
                    if (++nend == kTeams) break;
                }
            }
        }
    } else if (warp == kFWarp) {
        // ============================ F: line index of every tile of this CTA ===============================
        int64_t prevA = a.tile_begin - 1;                            // this CTA's previous tile
        uint64_t incl = a.tile_begin > 0 ? a.ws[kWsCarry] : 0ull;    // lines in all tiles up to and including tileB's predecessor chain
        uint64_t vB[kFK];
        int64_t tileB = -1, prevB = -1;
        uint32_t slotB = 0, parB = 0;
        bool haveB = false;
        unsigned long long f_complete = 0;
        int ends = 0;
        for (uint32_t seq = 0;; ++seq) {
            const uint32_t slot = seq & (kNS - 1), par = (seq / kNS) & 1u;
            // ---- A(seq): which tile, and the loads of every count published between it and its predecessor
            mbar_wait(bar_full + 8 * slot, par);
            const int32_t tileA = (int32_t)s_desc[slot * kDescWords + kDTile];
            uint64_t vA[kFK];
#pragma unroll
            for (int i = 0; i < kFK; ++i) {
                const int64_t idx = prevA + 1 + lane + 32 * i;
                vA[i] = (tileA >= 0 && idx < (int64_t)tileA) ? ld_relaxed(tile_state + idx) : kFlagAgg;
            }
            // ---- B(seq - 1): finish the previous tile (its loads had a whole tile period to land)
            if (haveB) {
                mbar_wait(bar_scanned + 8 * slotB, parB);
                const uint32_t count = s_desc[slotB * kDescWords + kDCount];
                uint64_t sum = 0;
#pragma unroll
                for (int i = 0; i < kFK; ++i) {
                    uint64_t v = vB[i];
                    const int64_t idx = prevB + 1 + lane + 32 * i;
                    while ((v >> 62) == 0) v = ld_relaxed(tile_state + idx);
                    sum += v & kValueMask;
                }
                for (int64_t idx = prevB + 1 + lane + 32 * kFK; idx < tileB; idx += 32) {   // rare: a long gap
                    uint64_t v = ld_relaxed(tile_state + idx);
                    while ((v >> 62) == 0) v = ld_relaxed(tile_state + idx);
                    sum += v & kValueMask;
                }
                const uint64_t base = incl + warp_sum_u64(sum);
                incl = base + count;
                const uint8_t *sp = s_slots + slotB * kSlot;
                const uint16_t *list = s_list + slotB * kNlCap;
                const size_t byte0 = (size_t)tileB * kTileBytes;
                const uint32_t base_phase = (uint32_t)base & pm;
                const int64_t q0 = (int64_t)(base >> ls);
                const uint32_t jr0 = (want - base_phase) & pm;       // first newline (rel) that precedes a field line
                // entry structure at the newlines before the first row's (one_line_buffer.py:155-173, fastq_buffer.py:38-45)
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
                    if (count > 0) {                                  // last complete entry of the tile
                        const uint32_t last = count - 1u;
                        const uint32_t back = (base_phase + last - pm) & pm;
                        if (last >= back && last - back < (uint32_t)kNlCap)
                            f_complete = max(f_complete, (unsigned long long)(byte0 + list[last - back] + 1));
                    }
                    if (tileB == 0 && a.n > 0 && sp[0] != a.header_char)
                        atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], 0ll);
                    if (tileB == a.n_tiles_total - 1) a.status[BNPK_ST_N_LINES] = (int64_t)(base + count);
                    if (tileB == (int64_t)tile_end - 1) a.ws[kWsCarry] = base + count;
                    *reinterpret_cast<volatile uint64_t *>(s_desc + slotB * kDescWords + kDBase) = base;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready + 8 * slotB);
            }
            if (tileA < 0) {
                if (lane == 0) mbar_arrive(bar_ready + 8 * slot);
                haveB = false;
                if (++ends == kTeams) break;
                continue;
            }
#pragma unroll
            for (int i = 0; i < kFK; ++i) vB[i] = vA[i];
            tileB = tileA; prevB = prevA; slotB = slot; parB = par; haveB = true;
            prevA = tileA;
        }
        if (lane == 0 && f_complete)
            atomicMax((unsigned long long *)&a.status[BNPK_ST_N_COMPLETE_BYTES], f_complete);
    } else if (warp >= kRW) {
        // ============================ S: newline masks, sorted newline list, tile count ======================
        const int sw = warp - kRW;
        ScanLane sl;
        sl.init(lane);
        int ends = 0;
        for (uint32_t seq = 0;; ++seq) {
            const uint32_t slot = seq & (kNS - 1), par = (seq / kNS) & 1u;
            mbar_wait(bar_full + 8 * slot, par);
            const int32_t tile = (int32_t)s_desc[slot * kDescWords + kDTile];
            if (tile < 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_scanned + 8 * slot);
                if (++ends == kTeams) break;
                continue;
            }
            uint8_t *sp = s_slots + slot * kSlot;
            const size_t byte0 = (size_t)tile * kTileBytes;
            const int staged = (int)min((size_t)kSlot, a.n - byte0);
            if (staged & 15) {                                        // the chunk's last bytes: not a multiple of 16
                const int t0 = staged & ~15;
                if (sw == 0 && lane < (staged & 15)) sp[t0 + lane] = a.chunk[byte0 + t0 + lane];
                scan_bar();
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
            if (lane == 0) s_wsum[(seq & 1u) * kSW + sw] = t0 + (tot >> 16);
            if (sw == kSW - 1) {                                      // first newline of the halo: end of the crossing row
                const int valid = min(max(staged - kTileBytes - 16 * lane, 0), 16);
                const uint32_t mm = newline_mask16(lds128(sp + kTileBytes + 16 * lane)) & ((1u << valid) - 1u);
                const unsigned b = __ballot_sync(0xffffffffu, mm != 0);
                const int srcl = b ? __ffs(b) - 1 : 0;
                const uint32_t pos = (uint32_t)(kTileBytes + 16 * lane + __ffs(mm) - 1);
                const uint32_t first = __shfl_sync(0xffffffffu, pos, srcl);
                if (lane == 0) s_desc[slot * kDescWords + kDCross] = b ? first : kNoCross;
            }
            scan_bar();                                               // warp totals of this tile visible (double-buffered)
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kSW; ++w) {
                const uint32_t v = s_wsum[(seq & 1u) * kSW + w];
                total += v;
                if (w < sw) before += v;
            }
            uint16_t *list = s_list + slot * kNlCap;
            const uint32_t pos0 = 4096u * (uint32_t)sw + 64u * (uint32_t)lane;
            emit_positions(nl0, before + (inc & 0xFFFFu) - cnt0, pos0, list, 0u);
            emit_positions(nl1, before + t0 + (inc >> 16) - cnt1, pos0 + 2048u, list, 0u);
            if (sw == 0 && lane == 0) {
                s_desc[slot * kDescWords + kDCount] = total;
                st_relaxed(tile_state + tile, kFlagAgg | (uint64_t)total);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_scanned + 8 * slot);
        }
    } else {
        // ============================ R: rows -> codes -> k-mers -> histogram =================================
        const int team = warp / kTeamWarps, tw = warp % kTeamWarps;
        const bool cr = a.status[BNPK_ST_CR] != 0;
        const uint64_t hmask = (a.n_bins & (a.n_bins - 1)) == 0 ? a.n_bins - 1 : 0;
        const uint64_t kmask = (1ull << (2 * a.k)) - 1;
        const bool fast = hmask && hmask <= 0x3FFFFFFFull;
        const uint32_t m32x4 = (uint32_t)(hmask & kmask) << 2;       // byte-offset mask into the table
        const uint32_t hist_sa = smem_addr(s_hist);
        uint32_t acc_bases = 0, acc_values = 0;                       // per lane: well inside 32 bits for any chunk
        unsigned long long last_start = 0, last_index = 0;             // 1 + start / entry of the last row this warp counted
        unsigned long long r_complete = 0;
        const uint32_t sub = (uint32_t)lane & 3u, q = (uint32_t)lane >> 2;
        const int src1 = (lane & ~3) | (int)((sub + 1u) & 3u), src2 = (lane & ~3) | (int)((sub + 2u) & 3u),
                  src3 = (lane & ~3) | (int)((sub + 3u) & 3u);
        const uint32_t cshift = 3u + ls, npc = 1u << cshift;          // newlines per chunk of 8 rows
        const int src_s = (int)(q << ls);                             // the lane that holds my row's start newline
        const uint32_t ph = (want + (uint32_t)lane) & pm;             // phase of my newline
        const bool chk_h = (uint32_t)lane < npc && ph == pm, chk_p = (uint32_t)lane < npc && a.check_plus && ph == 1u;

        auto defer_row = [&](uint64_t start, uint64_t r) {
            const unsigned long long d = atomicAdd((unsigned long long *)(a.ws + kWsDeferred), 1ull);
            if (d < a.deferred_cap) {
                a.deferred[2 * d] = start;
                a.deferred[2 * d + 1] = r;
            } else {
                a.status[BNPK_ST_OVERFLOW] = 1;
            }
        };

        for (uint32_t seq = (uint32_t)team;; seq += kTeams) {
            const uint32_t slot = seq & (kNS - 1), par = (seq / kNS) & 1u;
            mbar_wait(bar_ready + 8 * slot, par);
            const int32_t tile = (int32_t)s_desc[slot * kDescWords + kDTile];
            if (tile < 0) break;
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
            const int n_chunks = (tile_nl > jr0) ? (int)((tile_nl - jr0 + npc - 1u) >> cshift) : 0;

            // one chunk: 8 rows = npc consecutive newlines of the list window that starts at newline index wb
            auto do_chunk = [&](int c, uint32_t wb) {
                const uint32_t gi = jr0 + ((uint32_t)c << cshift) + (uint32_t)lane;   // my newline (tile-relative index)
                uint32_t p = 0xFFFFu;
                if ((uint32_t)lane < npc && gi < tile_nl) p = list[gi - wb];
                if ((chk_h || chk_p) && p != 0xFFFFu && byte0 + p + 1 < a.n) {
                    const uint32_t ch = sp[p + 1];
                    if (chk_h && ch != a.header_char)
                        atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], (long long)(q0 + ((base_phase + gi + 1u) >> ls)));
                    if (chk_p && ch != '+')
                        atomicMin((long long *)&a.status[BNPK_ST_BAD_PLUS_ENTRY], (long long)(q0 + ((base_phase + gi) >> ls)));
                }
                const uint32_t ps = __shfl_sync(0xffffffffu, p, src_s), pe = __shfl_sync(0xffffffffu, p, src_s + 1);
                const int s = 8 * c + (int)q;                         // my row (tile-relative)
                bool act = ps != 0xFFFFu;
                int b0 = (int)ps + 1, e = (int)pe;
                if (act && pe == 0xFFFFu) {                           // the row ends beyond the tile proper
                    if (crossM != kNoCross) {
                        e = (int)crossM;
                    } else {                                          // not terminated inside the slot
                        if (sub == 0 && byte0 + staged < a.n) defer_row(byte0 + b0, (uint64_t)(r_first + s));
                        act = false;                                  // (else: unterminated last line, not an entry)
                    }
                }
                if (act && cr && e > b0 && sp[e - 1] == '\r') e -= 1;
                if (act && e - b0 > kRowMax) {
                    if (sub == 0) defer_row(byte0 + b0, (uint64_t)(r_first + s));
                    act = false;
                }
                const int L = act ? e - b0 : 0;
                const int npos = max(L - a.k + 1, 0);
                if (act && sub == 0) {
                    acc_bases += (uint32_t)L;
                    acc_values += (uint32_t)npos;
                    if (s == n_rows_tile - 1) {                       // the tile's last counted row (see uncount_kernel)
                        last_start = max(last_start, (unsigned long long)(byte0 + b0) + 1ull);
                        last_index = max(last_index, (unsigned long long)(r_first + s) + 1ull);
                    }
                }
                const int A0 = b0 >> 4, A1 = (e - 1) >> 4;
                const uint32_t o = (uint32_t)b0 & 15u;
                const int my_rounds = L > 0 ? ((A1 - A0 + 1) + 3) >> 2 : 0;
                const int R = __reduce_max_sync(0xffffffffu, my_rounds);
                auto enc = [&](int r) -> uint32_t {
                    const int u = A0 + 4 * r + (int)sub;
                    if (L <= 0 || u > A1) return 0u;
                    const uint4 qq = lds128(sp + 16 * u);
                    const int lo = max(b0 - 16 * u, 0), hi = min(e - 16 * u, 16);
                    const uint32_t seq16 = (0xFFFFu >> (16 - hi)) & (0xFFFFu << lo);
                    uint32_t bad;
                    const uint32_t codes = encode_unit<ENC>(qq, seq16, s_lut, bad);
                    if (bad) {                                        // rare: exact position, byte by byte
                        for (int pp = 16 * u + lo; pp < 16 * u + hi; ++pp) {
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
                    }
                    return codes;
                };
                uint32_t c_cur = enc(0);
                for (int r = 0; r < R; ++r) {
                    const uint32_t c_nxt = enc(r + 1);
                    // code words of the next aligned units of my row: lanes of my quad, this round or the next
                    const uint32_t x1 = __shfl_sync(0xffffffffu, c_cur, src1), y1 = __shfl_sync(0xffffffffu, c_nxt, src1);
                    const uint32_t x2 = __shfl_sync(0xffffffffu, c_cur, src2), y2 = __shfl_sync(0xffffffffu, c_nxt, src2);
                    const uint32_t w1 = sub + 1u >= 4u ? y1 : x1, w2 = sub + 2u >= 4u ? y2 : x2;
                    const int left = npos - 16 * (4 * r + (int)sub);  // k-mers that start in my block of 16 bases
                    if (fast) {
                        if (__any_sync(0xffffffffu, left > 0)) {
                            // stream pre-shifted left by two bits: (window & mask) is the table's byte offset
                            const bool z = o == 0u;
                            const uint32_t p0 = z ? 0u : c_cur, p1 = z ? c_cur : w1, p2 = z ? w1 : w2;
                            const uint32_t sh = (2u * o + 30u) & 31u;
                            const uint32_t a0 = __funnelshift_r(p0, p1, sh), a1 = __funnelshift_r(p1, p2, sh);
                            if constexpr (SMEM_HIST) {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {      // two half blocks of eight
                                    if (__all_sync(0xffffffffu, left >= 8 * hb + 8)) {   // every lane has all eight
#pragma unroll
                                        for (int t = 8 * hb; t < 8 * hb + 8; ++t)
                                            hist_inc(hist_sa + ((t == 0 ? a0 : __funnelshift_r(a0, a1, 2 * t)) & m32x4));
                                    } else if (__any_sync(0xffffffffu, left > 8 * hb)) {
#pragma unroll
                                        for (int t = 8 * hb; t < 8 * hb + 8; ++t)
                                            hist_add_val(hist_sa + ((t == 0 ? a0 : __funnelshift_r(a0, a1, 2 * t)) & m32x4), (uint32_t)(t - left) >> 31);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int t = 0; t < 16; ++t) {
                                    const uint32_t v = (t == 0 ? a0 : __funnelshift_r(a0, a1, 2 * t)) & m32x4;
                                    if (t < left) {
                                        if constexpr (HIST == 2) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.hist32) + v), 1u);
                                        else atomicAdd(reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(a.hist) + 2 * (size_t)v), 1ull);
                                    }
                                }
                            }
                        }
                    } else {
                        const uint32_t x3 = __shfl_sync(0xffffffffu, c_cur, src3), y3 = __shfl_sync(0xffffffffu, c_nxt, src3);
                        const uint32_t w3 = sub + 3u >= 4u ? y3 : x3;
                        if (left > 0) {
                            const uint32_t sh = 2u * o;
                            const uint32_t a0 = __funnelshift_r(c_cur, w1, sh), a1 = __funnelshift_r(w1, w2, sh),
                                           a2 = __funnelshift_r(w2, w3, sh);
#pragma unroll 4
                            for (int t = 0; t < 16; ++t) {
                                if (t < left) {
                                    const uint32_t lo32 = __funnelshift_r(a0, a1, 2 * t), hi32 = __funnelshift_r(a1, a2, 2 * t);
                                    const uint64_t h = (((uint64_t)hi32 << 32) | lo32) & kmask;
                                    const uint64_t b = hmask ? (h & hmask) : (h % a.n_bins);
                                    if constexpr (HIST == 2) atomicAdd(a.hist32 + b, 1u);
                                    else if constexpr (HIST == 1) atomicAdd(s_hist + (uint32_t)b, 1u);
                                    else atomicAdd(a.hist + b, 1ull);
                                }
                            }
                        }
                    }
                    c_cur = c_nxt;
                }
            };

            if (tile_nl <= (uint32_t)kNlCap) {
                // chunks go round the team's warps; the start rotates so that the odd chunk does not always hit the same warp
                const int first = (tw + kTeamWarps - (int)((seq / kTeams) % kTeamWarps)) % kTeamWarps;
                for (int c = first; c < n_chunks; c += kTeamWarps) do_chunk(c, 0u);
            } else if (tw == 0) {
                // rare: more newlines than the list holds.  One warp walks the tile in windows of the list, which it
                // rebuilds itself from the bytes (window 0 is what the scan warps left).
                ScanLane sl;
                sl.init(lane);
                const int cpw = kWinStep >> cshift;                   // chunks per window
                for (int wk = 0; wk * cpw < n_chunks; ++wk) {
                    const uint32_t wb = (uint32_t)(wk * kWinStep);
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
                    if (lane == 0) {                                  // last complete entry of the tile, if this window holds it
                        const uint32_t last = tile_nl - 1u;
                        const uint32_t back = (base_phase + last - pm) & pm;
                        if (last >= back && last - back - wb < (uint32_t)kNlCap && last - back >= wb)
                            r_complete = max(r_complete, (unsigned long long)(byte0 + list[last - back - wb] + 1));
                    }
                    for (int c = wk * cpw; c < min(n_chunks, (wk + 1) * cpw); ++c) do_chunk(c, wb);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8 * slot);
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
    const size_t smem = (size_t)kFixedBytes + (HIST == 1 ? ((a.n_bins * 4 + 127) & ~(uint64_t)127) : 0);
    BNPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFixedBytes + kMaxBins * 4));
    const int64_t n_tiles = a.tile_end - a.tile_begin;
    if (n_tiles <= 0) return 0;
    const int64_t grid = std::min<int64_t>(n_tiles, (int64_t)sm_count());
    profile_before(st);
    kern<<<(unsigned)grid, kCta, smem, st>>>(a);
    profile_after(st);
    BNPK_LAUNCHED("tile_ws_kernel");
    return 0;
}

template <int ENC>
static int launch_enc(const TileArgs &a, bool smem_hist, cudaStream_t st) {
    if (smem_hist) return launch_t<ENC, 1>(a, st);
    return a.hist32 ? launch_t<ENC, 2>(a, st) : launch_t<ENC, 0>(a, st);
}

}  // namespace ws

int launch_ws_count(const TileArgs &a, int enc_mode, bool smem_hist, cudaStream_t st) {
    switch (enc_mode) {
        case BNPK_ENC_ASCII_ACGT: return ws::launch_enc<BNPK_ENC_ASCII_ACGT>(a, smem_hist, st);
        case BNPK_ENC_ASCII_ACTG: return ws::launch_enc<BNPK_ENC_ASCII_ACTG>(a, smem_hist, st);
        case BNPK_ENC_CODES: return ws::launch_enc<BNPK_ENC_CODES>(a, smem_hist, st);
        case BNPK_ENC_LUT: return ws::launch_enc<BNPK_ENC_LUT>(a, smem_hist, st);
    }
    return set_err(BNPK_E_BADARG, "bad enc_mode");
}

}  // namespace bnpk
