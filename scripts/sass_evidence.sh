#!/bin/bash
# Development aid: which sm_100a instructions the built library really uses, per kernel family -- counts of the mnemonics that
# matter for the design claims (TMA bulk copies, mbarrier, cluster barrier / distributed shared memory, packed dot products,
# fused add-min-max, warp votes / reductions).  usage: scripts/sass_evidence.sh > profiles/r02_sass_evidence.txt
lib=codec_tcc_b200/lib/libpeeb200.so
for fn in pee2_count_kernelItLi256ELi3 pee2_embed_kernelItLi256ELi3 pee2_extract_kernelItLi256ELi3 pee2_hist_kernelItLi256ELi3 \
          pee2_cluster_embed_kernelItLi512 pee2_cluster_extract_kernelItLi512 med_extract_kernelItLb1 pbr_pack_kernelILb0ELb1 \
          pbr_expand_kernelILb0ELb1 moments_f64_kernel; do
  sym=$(cuobjdump -elf $lib 2>/dev/null | grep -o "_ZN4peeb[0-9]*${fn}[A-Za-z0-9_]*" | grep -v "_param" | grep -v "[$]" | sort -u | head -1)
  [ -z "$sym" ] && continue
  echo "== $sym"
  cuobjdump -sass -fun "$sym" $lib 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' > /tmp/sass_ops.txt
  echo "   instructions: $(wc -l < /tmp/sass_ops.txt)"
  for pat in 'UBLKCP' 'SYNCS' 'UCGABAR_ARV' 'UCGABAR_WAIT' 'MAPA' 'ST\.E\.|^ST$|^ST\.' 'IDP\.2A' 'IDP\.4A' 'VIADDMNMX' 'VIMNMX' 'VIMNMX3' 'REDUX' 'VOTE' 'MATCH' 'SHFL' 'ATOMS' 'REDG|RED\.' 'LDS\.128' 'STS\.128' 'LDG\.E\.128|LDG.*128' 'POPC' 'DFMA|DADD|DMUL' 'BAR'; do
    c=$(grep -cE "^($pat)" /tmp/sass_ops.txt)
    [ "$c" -gt 0 ] && printf "   %-28s %6d\n" "$pat" "$c"
  done
done
