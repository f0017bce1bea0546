#!/bin/bash
# Per-kernel SASS evidence for profiles/: for every kernel in libgcnk.so, the counts of the mnemonics that show which
# hardware path it uses (tcgen05 = UTCHMMA/UTCQMMA..., TMA = UTMALDG/UBLKCP, mbarrier = SYNCS, mma.sync = HMMA/IMMA,
# 128-bit loads = LDG.E.128).  Usage: tools/sass_summary.sh > profiles/rNN_sass_summary.txt
set -e
lib=${1:-cuda_gcn_b200/libgcnk.so}
echo "# $(basename $lib) sha256 $(sha256sum $lib | cut -c1-16); cuobjdump -sass, sm_100a; columns: kernel | UTC*MMA (tcgen05.mma) | UTCBAR/UTCCP | LDTM/STTM (tcgen05.ld/st) | UTMALDG/UTMASTG (TMA tensor) | UBLKCP (bulk copy) | SYNCS (mbarrier) | HMMA (mma.sync) | LDG.*128 | ST*.128 | total instr"
cuobjdump -sass "$lib" | awk '
/Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; tot=0; next }
/^ +\/\*[0-9a-f]{4}\*\// { tot++; s=$0;
  if (s ~ /UTC[A-Z]*MMA/) c["mma"]++; if (s ~ /UTCBAR|UTCCP/) c["bar"]++; if (s ~ /LDTM|STTM/) c["tm"]++;
  if (s ~ /UTMALDG|UTMASTG|UTMAPF/) c["tma"]++; if (s ~ /UBLKCP/) c["blk"]++; if (s ~ /SYNCS/) c["syn"]++;
  if (s ~ /[ .]HMMA|[ .]IMMA/) c["hmma"]++; if (s ~ /LDG\.[A-Z.]*128/) c["ldg"]++; if (s ~ /ST[GS]\.[A-Z.]*128/) c["st"]++; }
END { emit() }
function emit() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", name, c["mma"], c["bar"], c["tm"], c["tma"], c["blk"], c["syn"], c["hmma"], c["ldg"], c["st"], tot }
' | c++filt | sort
