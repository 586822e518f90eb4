// tile_kernels.cu -- single-pass kernels over raw chunk bytes (K1 line split, K6 fused count).
//
// One persistent CTA per resident slot; tiles of kTileBytes are handed out in order by an
// atomic ticket.  Per tile: coalesced uint4 streaming loads -> in-register byte->(2-bit code,
// newline, valid) transform -> 8 bytes of shared memory per 16 input bytes; per-tile newline
// count -> decoupled look-back -> global line index of every newline (the row-offset vector
// lives in shared memory only); one warp per read row: rolling 2-bit hash read straight from
// the packed stream, optional warp-shuffle sliding minimum, privatised shared-memory histogram
// (or global atomics for big tables).
#include "bnpk_host.h"

namespace bnpk {

struct TileArgs {
    const uint8_t *chunk;
    size_t n;
    int64_t tile_begin, tile_end;  // tiles handled by this launch
    int lpe, field_line, start_offset;
    uint32_t header_char;
    int check_plus;
    int trim_cr;                   // -1 auto (status[CR]), 0, 1
    int64_t *status;
    uint64_t *ws;                  // header | tile_state[] | deferred[]
    int64_t n_tiles_total;
    size_t deferred_cap;
    // split
    int64_t *starts;
    int32_t *lens;
    size_t max_rows;
    // count
    const uint8_t *lut;
    int k, window;                 // window = 0: k-mers; else minimizers over `window` bases
    uint64_t n_bins;
    unsigned long long *hist;
};

__device__ __forceinline__ uint64_t *ws_tile_state(uint64_t *ws) { return ws + kWsHeaderWords; }
__device__ __forceinline__ uint64_t *ws_deferred(uint64_t *ws, int64_t n_tiles_total) {
    return ws + kWsHeaderWords + n_tiles_total;
}

// -------------------------------------------------------------------------------------------
// init: decide '\r' trimming like OneLineBuffer._modify_for_carriage_return
// (io/one_line_buffer.py:175-182): trim iff the header line of one of the first
// `lines_per_entry` entries ends in '\r'.
// -------------------------------------------------------------------------------------------
__global__ void cr_detect_kernel(const uint8_t *chunk, size_t n, int lpe, int trim_cr, int64_t *status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t cr = 0;
    if (trim_cr == 1) cr = 1;
    if (trim_cr < 0) {
        size_t p = 0;
        int line = 0;
        const int max_lines = lpe * lpe;
        const size_t limit = n < (size_t)(8u << 20) ? n : (size_t)(8u << 20);
        for (; p < limit && line < max_lines; ++p) {
            if (chunk[p] == '\n') {
                if (line % lpe == 0 && p > 0 && chunk[p - 1] == '\r') { cr = 1; break; }
                ++line;
            }
        }
    }
    status[BNPK_ST_CR] = cr;
}

// -------------------------------------------------------------------------------------------
// block-wide exclusive scan of one uint32 per thread (kTileThreads threads)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kTileWarps ? s_warp[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kTileWarps) s_warp[lane] = winc - w;
        if (lane == kTileWarps - 1) s_warp[kTileWarps] = winc;
    }
    __syncthreads();
    total = s_warp[kTileWarps];
    const uint32_t r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

// -------------------------------------------------------------------------------------------
// the tile kernel.  MODE 0 = split (write starts/lens), MODE 1 = fused count.
// -------------------------------------------------------------------------------------------
template <int MODE, int ENC, bool SMEM_HIST, bool MINIMIZER>
__global__ void __launch_bounds__(kTileThreads, MODE == 0 ? 2 : 1) tile_kernel(const TileArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_codes = smem;                                  // kStagedUnits + 4
    uint32_t *s_flags = s_codes + kStagedUnits + 4;            // kStagedUnits
    uint32_t *s_warp = s_flags + kStagedUnits;                 // kTileWarps + 1 (+pad to 32)
    uint32_t *s_misc = s_warp + 32;                            // 16 words
    uint16_t *s_rows = reinterpret_cast<uint16_t *>(s_misc + 16);                   // kRowCap
    uint8_t *s_lut = reinterpret_cast<uint8_t *>(s_rows + kRowCap);                 // 256
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_lut + 256);                   // n_bins (SMEM_HIST)
    __shared__ int64_t s_line_base;
    __shared__ int64_t s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *tile_state = ws_tile_state(a.ws);
    const bool cr = (MODE == 0 || a.field_line >= 0) && (a.status[BNPK_ST_CR] != 0);

    if (MODE == 1) {
        if (ENC == BNPK_ENC_LUT && tid < 256) s_lut[tid] = a.lut[tid];
        if (SMEM_HIST)
            for (uint32_t b = tid; b < a.n_bins; b += kTileThreads) s_hist[b] = 0;
    }
    if (tid < 4) s_codes[kStagedUnits + tid] = 0;
    HistTarget ht;
    ht.global = a.hist;
    ht.smem = s_hist;
    ht.n_bins = a.n_bins;
    ht.mask = (a.n_bins & (a.n_bins - 1)) == 0 ? a.n_bins - 1 : 0;
    ht.delta = 1ull;
    uint64_t acc_bases = 0, acc_values = 0;     // per-thread statistics, flushed once
    __syncthreads();

    while (true) {
        if (tid == 0) s_ticket = a.tile_begin + (int64_t)atomicAdd((unsigned long long *)(a.ws + kWsTicket), 1ull);
        __syncthreads();
        const int64_t tile = s_ticket;
        if (tile >= a.tile_end) break;
        const size_t byte0 = (size_t)tile * kTileBytes;
        const int tile_len = (int)min((size_t)kTileBytes, a.n - byte0);
        const int staged_len = (MODE == 1) ? (int)min((size_t)(kTileBytes + kHaloBytes), a.n - byte0) : tile_len;
        const int n_units = (staged_len + 15) >> 4;

        // ---- stage: global -> registers -> (codes, flags) in shared memory -----------------
        {
            const bool aligned = ((reinterpret_cast<uintptr_t>(a.chunk) & 15) == 0);
            constexpr int kMaxPer = (kStagedUnits + kTileThreads - 1) / kTileThreads;  // 5
            uint4 q[kMaxPer];
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int u = tid + j * kTileThreads;
                if (u < n_units) {
                    if (aligned && (u + 1) * 16 <= staged_len)
                        q[j] = ld_stream(reinterpret_cast<const uint4 *>(a.chunk + byte0) + u);
                    else
                        q[j] = load_unit_guarded(a.chunk, a.n, (int64_t)byte0 + (int64_t)u * 16);
                }
            }
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int u = tid + j * kTileThreads;
                if (u < kStagedUnits) {
                    uint32_t c = 0, f = 0;
                    if (u < n_units) {
                        if (MODE == 1) encode_unit<ENC>(q[j], s_lut, c, f);
                        else {
                            f = bytes_lsb_to_nibble(__vcmpeq4(q[j].x, 0x0A0A0A0Au)) |
                                (bytes_lsb_to_nibble(__vcmpeq4(q[j].y, 0x0A0A0A0Au)) << 4) |
                                (bytes_lsb_to_nibble(__vcmpeq4(q[j].z, 0x0A0A0A0Au)) << 8) |
                                (bytes_lsb_to_nibble(__vcmpeq4(q[j].w, 0x0A0A0A0Au)) << 12);
                        }
                        // bytes past the end of the data are neither newline nor valid
                        const int over = (u + 1) * 16 - staged_len;
                        if (over > 0) { const uint32_t keep = 0xFFFFu >> over; f &= keep | (keep << 16); }
                    }
                    s_codes[u] = c;
                    s_flags[u] = f;
                }
            }
        }
        __syncthreads();

        // ---- newline census of the tile proper (not the halo) -----------------------------
        // thread t owns units [4t, 4t+4) = 64 bytes
        uint64_t nlmask = 0;
        {
            const uint4 f4 = *reinterpret_cast<const uint4 *>(s_flags + 4 * tid);
            nlmask = (uint64_t)(f4.x & 0xFFFFu) | ((uint64_t)(f4.y & 0xFFFFu) << 16) |
                     ((uint64_t)(f4.z & 0xFFFFu) << 32) | ((uint64_t)(f4.w & 0xFFFFu) << 48);
            const int first_byte = tid * 64;
            if (first_byte >= tile_len) nlmask = 0;
            else if (first_byte + 64 > tile_len) nlmask &= (~0ull) >> (64 - (tile_len - first_byte));
        }
        uint32_t tile_nl;
        const uint32_t my_excl = block_excl_scan((uint32_t)__popcll(nlmask), s_warp, tile_nl);
        if (warp == 0) {
            const uint64_t excl = lookback_exclusive(tile_state, tile, tile_nl, lane);
            if (lane == 0) s_line_base = (int64_t)excl;
        }
        if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0; s_misc[2] = 0; }
        __syncthreads();
        const int64_t line_base = s_line_base;
        const int lpe = a.lpe;

        // rows whose field line starts in this tile: the newline j with (j+1) % lpe == field_line
        // ends the previous line; the field line starts right after it.
        const int64_t want = ((a.field_line - 1) % lpe + lpe) % lpe;   // j % lpe we look for
        const int64_t j0 = line_base + (((want - line_base) % lpe) + lpe) % lpe;
        const int64_t r_first = (j0 + 1) / lpe;
        const int64_t n_rows_tile = (line_base + (int64_t)tile_nl - 1 >= j0)
                                        ? (line_base + (int64_t)tile_nl - 1 - j0) / lpe + 1 : 0;

        // ---- per-newline events ---------------------------------------------------------------
        uint32_t my_complete = 0;   // tile-relative (p+1) of the last record-ending newline I own
        auto newline_events = [&](int round) {
            uint64_t m = nlmask;
            int64_t j = line_base + my_excl;
            while (m) {
                const int bit = __ffsll((long long)m) - 1;
                m &= m - 1;
                const int p = tid * 64 + bit;                      // tile-relative newline position
                const size_t gp = byte0 + p;                       // global newline position
                const int phase = (int)(j % lpe);
                if (round == 0) {
                    if (phase == lpe - 1) {
                        my_complete = p + 1;
                        // next entry's header char (one_line_buffer.py:155-173)
                        if (gp + 1 < a.n && a.chunk[gp + 1] != a.header_char)
                            atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], (long long)((j + 1) / lpe));
                    }
                    if (a.check_plus && phase == 1) {                 // fastq_buffer.py:38-45
                        if (gp + 1 < a.n && a.chunk[gp + 1] != '+')
                            atomicMin((long long *)&a.status[BNPK_ST_BAD_PLUS_ENTRY], (long long)(j / lpe));
                    }
                }
                if (MODE == 0) {
                    // split: start and end of the wanted line are published independently;
                    // lens[r] accumulates (end - start) mod 2^32 from two atomics.
                    if (phase == want) {
                        const int64_t r = (j + 1) / lpe;
                        if ((size_t)r < a.max_rows) {
                            const int64_t s = (int64_t)gp + 1 + a.start_offset;
                            a.starts[r] = s;
                            atomicSub((unsigned int *)&a.lens[r], (unsigned int)(uint64_t)s);
                        }
                    }
                    if (phase == a.field_line) {
                        const int64_t r = j / lpe;
                        if ((size_t)r < a.max_rows) {
                            int64_t e = (int64_t)gp;
                            if (cr && gp > 0 && a.chunk[gp - 1] == '\r') e -= 1;
                            atomicAdd((unsigned int *)&a.lens[r], (unsigned int)(uint64_t)e);
                        }
                    }
                } else {
                    if (phase == want) {
                        const int64_t slot = (j + 1) / lpe - r_first - (int64_t)round * kRowCap;
                        if (slot >= 0 && slot < kRowCap) s_rows[slot] = (uint16_t)(p + 1 + a.start_offset);
                    }
                }
                ++j;
            }
        };

        if (MODE == 0) {
            newline_events(0);
            if (tile == 0 && tid == 0) {
                if (a.n > 0 && a.chunk[0] != a.header_char)
                    atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], 0ll);
                if (a.field_line == 0 && a.max_rows > 0) {        // first line has no newline before it
                    a.starts[0] = a.start_offset;
                    atomicSub((unsigned int *)&a.lens[0], (unsigned int)a.start_offset);
                }
            }
            if (my_complete) atomicMax(&s_misc[0], my_complete);
            __syncthreads();
        } else {
            if (tile == 0 && tid == 0 && a.n > 0 && a.chunk[0] != a.header_char)
                atomicMin((long long *)&a.status[BNPK_ST_BAD_HEADER_ENTRY], 0ll);
            const int n_rounds = (int)((n_rows_tile + kRowCap - 1) / kRowCap);
            for (int round = 0; round < (n_rounds > 0 ? n_rounds : 1); ++round) {
                newline_events(round);
                if (round == 0 && my_complete) atomicMax(&s_misc[0], my_complete);
                __syncthreads();
                const int rows_here = (int)min((int64_t)kRowCap, n_rows_tile - (int64_t)round * kRowCap);
                for (int slot = warp; slot < rows_here; slot += kTileWarps) {
                    const int b0 = s_rows[slot];
                    const int64_t r = r_first + (int64_t)round * kRowCap + slot;
                    if (b0 > staged_len) continue;                  // start_offset ran past the data
                    const int e = find_newline(s_flags, b0, staged_len, lane);
                    if (e < 0) {
                        if (byte0 + staged_len >= a.n) continue;    // unterminated last line: not an entry
                        if (lane == 0) {                            // long row: defer
                            const unsigned long long d = atomicAdd((unsigned long long *)(a.ws + kWsDeferred), 1ull);
                            if (d < a.deferred_cap) {
                                uint64_t *def = ws_deferred(a.ws, a.n_tiles_total);
                                def[2 * d] = byte0 + b0;
                                def[2 * d + 1] = (uint64_t)r;
                            }
                        }
                        continue;
                    }
                    int L = e - b0;
                    if (cr && L > 0 && a.chunk[byte0 + e - 1] == '\r') L -= 1;
                    const int bad = find_invalid(s_flags, b0, b0 + L, lane);
                    if (bad >= 0) {
                        if (lane == 0)
                            atomicMin((long long *)&a.status[BNPK_ST_BAD_BASE], (long long)((r << 32) | (int64_t)(bad - b0)));
                        continue;
                    }
                    if (lane == 0) {
                        acc_bases += (uint64_t)L;
                        atomicMax(&s_misc[1], (uint32_t)b0 + 1u);
                        atomicMax(&s_misc[2], (uint32_t)(r - r_first) + 1u);
                    }
                    const int span = MINIMIZER ? a.window : a.k;
                    if (L >= span) acc_values += row_count<SMEM_HIST, MINIMIZER>(s_codes, b0, L, a.k, a.window, ht, lane);
                }
                __syncthreads();
            }
        }
        // ---- per-tile global bookkeeping (one atomic each) --------------------------------------
        if (tid == 0) {
            if (s_misc[0]) atomicMax((unsigned long long *)&a.status[BNPK_ST_N_COMPLETE_BYTES], (unsigned long long)(byte0 + s_misc[0]));
            if (MODE == 1 && s_misc[1]) {
                atomicMax((unsigned long long *)&a.status[BNPK_ST_LAST_ROW_START], (unsigned long long)(byte0 + s_misc[1] - 1) + 1ull);
                atomicMax((unsigned long long *)&a.status[BNPK_ST_LAST_ROW_INDEX], (unsigned long long)(r_first + s_misc[2] - 1) + 1ull);
            }
            if (tile == a.n_tiles_total - 1) a.status[BNPK_ST_N_LINES] = line_base + tile_nl;
        }
        __syncthreads();
    }

    // ---- flush ---------------------------------------------------------------------------------
    if (MODE == 1) {
        if (SMEM_HIST) {
            __syncthreads();
            for (uint32_t b = tid; b < a.n_bins; b += kTileThreads) {
                const uint32_t c = s_hist[b];
                if (c) atomicAdd(a.hist + b, (unsigned long long)c);
            }
        }
        acc_bases = warp_sum_u64(acc_bases);
        acc_values = warp_sum_u64(acc_values);
        if (lane == 0) {
            if (acc_bases) atomicAdd((unsigned long long *)&a.status[BNPK_ST_N_BASES], acc_bases);
            if (acc_values) atomicAdd((unsigned long long *)&a.status[BNPK_ST_N_VALUES], acc_values);
        }
    }
}

// n_records and friends once every tile is done
__global__ void finalize_status_kernel(int64_t *status, int lpe) {
    if (threadIdx.x == 0 && blockIdx.x == 0) status[BNPK_ST_N_RECORDS] = status[BNPK_ST_N_LINES] / lpe;
}

static size_t tile_smem_bytes(int mode, uint64_t n_bins, bool smem_hist) {
    size_t words = (kStagedUnits + 4) + kStagedUnits + 32 + 16;
    size_t bytes = words * 4 + kRowCap * 2 + 256;
    if (mode == 1 && smem_hist) bytes += n_bins * 4;
    return bytes;
}

template <int MODE, int ENC, bool SMEM_HIST, bool MINIMIZER>
static int launch_tile(const TileArgs &a, cudaStream_t st) {
    auto kern = tile_kernel<MODE, ENC, SMEM_HIST, MINIMIZER>;
    const size_t smem = tile_smem_bytes(MODE, a.n_bins, SMEM_HIST);
    static thread_local bool attr_done = false;  // per instantiation
    if (!attr_done) {
        BNPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    int per_sm = 1;
    BNPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem));
    if (per_sm < 1) return set_err(BNPK_E_BINS, "tile kernel does not fit shared memory");
    const int64_t n_tiles = a.tile_end - a.tile_begin;
    if (n_tiles <= 0) return 0;
    const int64_t grid = std::min<int64_t>(n_tiles, (int64_t)sm_count() * per_sm);
    profile_before(st);
    kern<<<(unsigned)grid, kTileThreads, smem, st>>>(a);
    profile_after(st);
    BNPK_LAUNCHED("tile_kernel");
    return 0;
}

template <int ENC>
static int launch_count_enc(const TileArgs &a, bool smem_hist, cudaStream_t st) {
    const bool mz = a.window > 0;
    if (smem_hist) return mz ? launch_tile<1, ENC, true, true>(a, st) : launch_tile<1, ENC, true, false>(a, st);
    return mz ? launch_tile<1, ENC, false, true>(a, st) : launch_tile<1, ENC, false, false>(a, st);
}

static int launch_count(const TileArgs &a, int enc_mode, bool smem_hist, cudaStream_t st) {
    switch (enc_mode) {
        case BNPK_ENC_ASCII_ACGT: return launch_count_enc<BNPK_ENC_ASCII_ACGT>(a, smem_hist, st);
        case BNPK_ENC_ASCII_ACTG: return launch_count_enc<BNPK_ENC_ASCII_ACTG>(a, smem_hist, st);
        case BNPK_ENC_CODES: return launch_count_enc<BNPK_ENC_CODES>(a, smem_hist, st);
        case BNPK_ENC_LUT: return launch_count_enc<BNPK_ENC_LUT>(a, smem_hist, st);
    }
    return set_err(BNPK_E_BADARG, "bad enc_mode");
}

size_t tile_workspace_bytes(size_t n) {
    const size_t n_tiles = (n + kTileBytes - 1) / kTileBytes + 1;
    const size_t deferred = n / kHaloBytes + 16;
    return (kWsHeaderWords + n_tiles + 2 * deferred) * sizeof(uint64_t);
}

bool use_smem_hist(int64_t n_bins, int hist_mode) {
    if (hist_mode == BNPK_HIST_GLOBAL) return false;
    return n_bins <= kSmemMaxBins;
}

int chunk_kmer_count_impl(const uint8_t *chunk, size_t n, size_t slice_begin, size_t slice_end, int final_slice,
                          int lpe, uint8_t header_char, int check_plus, int trim_cr, int enc_mode,
                          const uint8_t *lut256, int k, int window, int64_t n_bins, int hist_mode, int64_t *hist,
                          int64_t *status, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (k < 1 || k > 31) return set_err(BNPK_E_K, "k must be larger than 0 and smaller than 32");
    if (window != 0 && window < k) return set_err(BNPK_E_WINDOW, "kmer size must be smaller than window size");
    if (n_bins < 1) return set_err(BNPK_E_BINS, "n_bins must be positive");
    if (hist_mode == BNPK_HIST_SMEM && n_bins > kSmemMaxBins) return set_err(BNPK_E_BINS, "too many bins for the shared-memory histogram");
    if (enc_mode == BNPK_ENC_LUT && !lut256) return set_err(BNPK_E_BADARG, "lut256 required");
    if (lpe < 2 || slice_end > n || slice_begin > slice_end) return set_err(BNPK_E_BADARG, "bad slice");
    if (workspace_bytes < tile_workspace_bytes(n)) return set_err(BNPK_E_WORKSPACE, "workspace too small");
    if (n == 0) return 0;
    const int64_t n_tiles_total = (int64_t)((n + kTileBytes - 1) / kTileBytes);
    // a tile is complete once its staged region [t*T, min((t+1)*T + H, n)) is resident
    auto tiles_done_at = [&](size_t resident) -> int64_t {
        if (resident >= n) return n_tiles_total;
        if (resident < (size_t)(kTileBytes + kHaloBytes)) return 0;
        return (int64_t)((resident - kHaloBytes) / kTileBytes);
    };
    TileArgs a{};
    a.chunk = chunk; a.n = n;
    a.tile_begin = tiles_done_at(slice_begin);
    a.tile_end = final_slice ? n_tiles_total : tiles_done_at(slice_end);
    a.lpe = lpe; a.field_line = 1; a.start_offset = 0; a.header_char = header_char; a.check_plus = check_plus;
    a.trim_cr = trim_cr; a.status = status; a.ws = (uint64_t *)workspace; a.n_tiles_total = n_tiles_total;
    a.deferred_cap = n / kHaloBytes + 16;
    a.lut = lut256; a.k = k; a.window = window; a.n_bins = (uint64_t)n_bins; a.hist = (unsigned long long *)hist;
    if (slice_begin == 0) {
        BNPK_CUDA(cudaMemsetAsync(workspace, 0, (kWsHeaderWords + n_tiles_total + 1) * sizeof(uint64_t), st));
        cr_detect_kernel<<<1, 32, 0, st>>>(chunk, std::min(n, slice_end), lpe, trim_cr, status);
        BNPK_LAUNCHED("cr_detect_kernel");
    }
    BNPK_CUDA(cudaMemsetAsync(a.ws + kWsTicket, 0, sizeof(uint64_t), st));
    const bool smem_hist = use_smem_hist(n_bins, hist_mode);
    int rc = launch_count(a, enc_mode, smem_hist, st);
    if (rc) return rc;
    if (final_slice) {
        finalize_status_kernel<<<1, 32, 0, st>>>(status, lpe);
        BNPK_LAUNCHED("finalize_status_kernel");
        rc = count_fixups_impl(chunk, n, lpe, enc_mode, lut256, k, window, n_bins, hist, status,
                               (uint64_t *)workspace + kWsDeferred,
                               (uint64_t *)workspace + kWsHeaderWords + n_tiles_total, a.deferred_cap, st);
    }
    return rc;
}

int line_split_impl(const uint8_t *chunk, size_t n, int lpe, int field_line, int start_offset, uint8_t header_char,
                    int check_plus, int trim_cr, int64_t *starts, int32_t *lens, size_t max_rows, int64_t *status,
                    void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (lpe < 1 || field_line < 0 || field_line >= lpe) return set_err(BNPK_E_BADARG, "bad line layout");
    if (workspace_bytes < tile_workspace_bytes(n)) return set_err(BNPK_E_WORKSPACE, "workspace too small");
    if (n == 0) return 0;
    TileArgs a{};
    a.chunk = chunk; a.n = n;
    a.n_tiles_total = (int64_t)((n + kTileBytes - 1) / kTileBytes);
    a.tile_begin = 0; a.tile_end = a.n_tiles_total;
    a.lpe = lpe; a.field_line = field_line; a.start_offset = start_offset; a.header_char = header_char;
    a.check_plus = check_plus; a.trim_cr = trim_cr; a.status = status; a.ws = (uint64_t *)workspace;
    a.starts = starts; a.lens = lens; a.max_rows = max_rows; a.n_bins = 1;
    if (max_rows) BNPK_CUDA(cudaMemsetAsync(lens, 0, max_rows * sizeof(int32_t), st));
    BNPK_CUDA(cudaMemsetAsync(workspace, 0, (kWsHeaderWords + a.n_tiles_total + 1) * sizeof(uint64_t), st));
    cr_detect_kernel<<<1, 32, 0, st>>>(chunk, n, lpe, trim_cr, status);
    BNPK_LAUNCHED("cr_detect_kernel");
    BNPK_CUDA(cudaMemsetAsync(a.ws + kWsTicket, 0, sizeof(uint64_t), st));
    int rc = launch_tile<0, BNPK_ENC_ASCII_ACGT, false, false>(a, st);
    if (rc) return rc;
    finalize_status_kernel<<<1, 32, 0, st>>>(status, lpe);
    BNPK_LAUNCHED("finalize_status_kernel");
    return 0;
}

}  // namespace bnpk
