#!/bin/bash
# development (needs a library built with NVCC_EXTRA=-DBNPK_WS_DEBUG_KNOBS): time the fused count with parts of the row warps' work switched off (BNPK_WS_DEBUG bits:
# 1 = every lane counts into its own bank, 2 = no k-mer blocks, 4 = no encoding, 8 = row warps only free the slot)
for d in 0 1 2 6 8; do echo -n "dbg=$d: "; BNPK_WS_DEBUG=$d timeout 200 python tools/time_count.py 2>&1 | tail -1 | cut -d'|' -f1; done
