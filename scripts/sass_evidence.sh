#!/bin/bash
# Regenerates profiles/sass/*.txt: SASS mnemonic histograms and the TMA / mbarrier /
# peer-store instruction lines of the flagship kernels (cuobjdump -sass on the in-tree .so).
set -e
SO=bigslice_b200/_bscuda.so
mkdir -p profiles/sass
dump() {  # $1 = mangled-name regex, $2 = output stem
  cuobjdump -sass $SO 2>/dev/null | awk -v pat="$1" '
    /Function :/ {keep = ($0 ~ pat)} keep {print}' > /tmp/sass_$2.txt
  {
    echo "# $2 — SASS evidence (cuobjdump -sass $SO, function ~ /$1/)"
    echo "lines: $(wc -l < /tmp/sass_$2.txt)"
    echo; echo "## mnemonic histogram (memory / sync / atomic classes)"
    grep -oE "\b(UBLKCP[.A-Z0-9_]*|UTMA[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|ATOMS[.A-Z0-9_]*|ATOMG[.A-Z0-9_]*|REDG[.A-Z0-9_]*|RED[.A-Z0-9_]*|LDG[.A-Z0-9_]*|STG[.A-Z0-9_]*|LDS[.A-Z0-9_]*|STS[.A-Z0-9_]*|MEMBAR[.A-Z0-9_]*|FENCE[.A-Z0-9_]*|MATCH[.A-Z0-9_]*|BAR[.A-Z0-9_]*|NANOSLEEP|CCTL[.A-Z0-9_]*|ERRBAR|UTC[A-Z0-9.]*MMA|HMMA[.A-Z0-9_]*)" /tmp/sass_$2.txt | sort | uniq -c | sort -rn
    echo; echo "## selected instruction lines"
    grep -nE "UBLKCP|SYNCS|ATOMG|REDG|\.SYS|MEMBAR|MATCH|VOTE|createpolicy|EVICT|LDG.*\.(EF|EL)|STG.*STRONG" /tmp/sass_$2.txt | head -60
  } > profiles/sass/$2.txt
}
dump "hash_combine_tma_kernelIxLi0" hash_combine_tma_i64_add
dump "xchg_scatter_kernel" xchg_scatter
dump "xchg_table_scatter_buf_kernelILi8" xchg_table_scatter_buf
dump "xchg_plan_kernel" xchg_plan
dump "xchg_wait_kernel" xchg_wait
dump "rs_scatter_kernel" radix_scatter
dump "table_compact_kernelIxE" table_compact_i64
dump "hash_column_kernel" hash_column
dump "partition_scatter_kernel" partition_scatter
dump "gather_rows_kernel" gather_rows
dump "xchg_table_scatter_kernelILi8" xchg_table_scatter_tiles
dump "xchg_table_count_kernel" xchg_table_count
dump "segreduce_kernelIxLi0" segreduce_i64_add
dump "topk_hist_kernel" topk_hist
dump "topk_select_kernel" topk_select
dump "compact_" compact_indices
dump "expand_indices_kernel" expand_indices
dump "crc32_blocks_kernel" crc32_blocks
dump "rs_global_hist_kernel" radix_global_hist
ls -la profiles/sass
