#!/bin/bash
# scripts/make_sass_evidence.sh — SASS evidence for the claims DESIGN.md makes about the shipped kernels, from the
# library that is actually loaded (seaweedfs_b200/libswec.so).  Needs no GPU.  Writes profiles/sass_*.txt.
set -e
cd "$(dirname "$0")/.."
LIB=seaweedfs_b200/libswec.so
OUT=profiles
mnemonics() {  # histogram of SASS mnemonics of one function
  grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn
}
dump() {  # $1 = demangled-name regex, $2 = output file, $3 = title
  local sym
  sym=$(cuobjdump -sass $LIB | grep -E "Function : " | sed 's/.*Function : //' | while read -r f; do echo "$f $(echo "$f" | c++filt)"; done | grep -E "$1" | head -1 | awk '{print $1}')
  { echo "# $3"; echo "# library: $LIB   arch: $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
    echo "# function: $(echo "$sym" | c++filt)"; echo "# resources: $(cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "$sym" | tail -1)"
    echo "# --- mnemonic histogram"; cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | mnemonics
    if [ "${4:-full}" = full ]; then echo "# --- full listing"; cuobjdump -sass -fun "$sym" $LIB 2>/dev/null
    else echo "# --- excerpt: lines matching $4 (with context)"; cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -E -B2 -A2 "$4" | cut -c1-140; fi; } > "$2"
  echo "$2: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' "$2") instructions"
}
dump 'rs10x4_encode<512, 2, true, false>'  $OUT/sass_rs10x4_encode_512x2_blocked.txt "RS(10,4) encode, 512 threads x 2 column slices, BLOCKED layout, boost-clock multiply-by-2 step (the default launch)"
dump 'rs10x4_encode<512, 2, true, true>'   $OUT/sass_rs10x4_encode_512x2_blocked_lowpower.txt "RS(10,4) encode, same shape, low-power multiply-by-2 step (PRMT sign mask; taken under sustained load)" 'LDG|STG' 
dump 'swec_aot_recon<swec_aot_boost::SwecAotRecon14>' $OUT/sass_aot_recon_worst_case.txt "AOT reconstruct kernel, data shards 0-3 lost (worst case, BASELINE configs[2])" 'LDG|STG' 
dump 'swec_aot_recon<swec_aot_boost::SwecAotRecon0>'  $OUT/sass_aot_recon_single_loss.txt "AOT reconstruct kernel, data shard 0 lost (one output row)" 'LDG|STG' 
dump 'swec_table_kernel<10>' $OUT/sass_table_kernel_k10.txt "shared-memory table kernel (cold matrices): TMA bulk copy of the nibble tables = UBLKCP + SYNCS" 'UBLKCP|SYNCS' 
{ echo "# sm_100a only?"; cuobjdump -lelf $LIB; echo "# vector memory ops over the whole library:"
  cuobjdump -sass $LIB | grep -oE 'LDG\.E[A-Z0-9.]*128[A-Z.]*|STG\.E[A-Z0-9.]*128|UBLKCP[A-Z0-9.]*|SYNCS[A-Z0-9.]*|HMMA|IMMA|UTCHMMA|UTCIMMA' | sed -E 's/\.[0-9A-Z]+$//;' | sort | uniq -c | sort -rn | head -20
  echo "# tensor-core instructions (expected none): $(cuobjdump -sass $LIB | grep -cE 'HMMA|IMMA|UTC.MMA' || true)"; } > $OUT/sass_library_summary.txt
cat $OUT/sass_library_summary.txt
