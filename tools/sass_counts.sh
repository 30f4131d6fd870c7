#!/bin/bash
# Regenerates profiles/r02_sass_counts.md: SASS mnemonic counts of the built library (no GPU needed).
SO=lemo_b200/_build/liblemo_b200.so
cuobjdump -sass $SO > /tmp/sass.txt 2>/dev/null
{
echo "# SASS evidence for the tcgen05 + TMA paths (round 2)"
echo
echo "Command: \`cuobjdump -sass lemo_b200/_build/liblemo_b200.so\` on the library built by \`__graft_entry__.build()\`"
echo "(nvcc 12.9, \`-gencode arch=compute_100a,code=sm_100a -lineinfo -O3\`).  Regenerate with \`tools/sass_counts.sh\`."
echo
echo "| SASS mnemonic | what it is | count |"
echo "|---|---|---:|"
row() { echo "| $1 | $2 | $(grep -c -E "$3" /tmp/sass.txt) |"; }
row UTCHMMA "tcgen05.mma (kind::f16 / kind::tf32), issued by one elected thread" "UTCHMMA"
row UTMALDG "TMA tensor load (cp.async.bulk.tensor)" "UTMALDG"
row LDTM "tcgen05.ld (TMEM -> registers, epilogues)" "LDTM"
row STTM "tcgen05.st (registers -> TMEM: weights-in-TMEM conv kernel)" "STTM"
row UTCBAR "tcgen05.commit -> mbarrier" "UTCBAR"
row "SYNCS" "mbarrier arrive / try_wait" "SYNCS"
row "UCGABAR" "cluster barrier (cluster split-K GEMM)" "UCGABAR"
row "warp-level HMMA / IMMA" "legacy mma.sync (none: every tensor-core op is tcgen05)" "[^C]HMMA|IMMA"
row FFMA "fp32 FMA (CUDA-core kernels)" "FFMA"
echo
echo "Kernels that contain tcgen05.mma (UTCHMMA), with their TMA-load and TMEM-load counts:"
echo
echo '```'
awk '/Function :/{f=$3} /UTCHMMA/{c[f]++} /UTMALDG/{t[f]++} /LDTM/{l[f]++} END{for(k in c) printf "%s UTCHMMA %d UTMALDG %d LDTM %d\n", k, c[k], t[k], l[k]}' /tmp/sass.txt | sort | c++filt | sed -E 's/\(.*\)//' | cut -c1-150
echo '```'
} > profiles/r02_sass_counts.md
