#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM / PDL / cluster use: tools/sass_evidence.sh > profiles/...
SO=${1:-byo-gan_b200/libbg_b200.so}
echo "SASS evidence, cuobjdump -sass $SO (sm_100a)."
echo "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, SYNCS = mbarrier ops,"
echo "ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents, REDG / RED / ATOMG = global reductions, LDGSTS = cp.async,"
echo "UCGABAR = barrier.cluster, MAPA = mapa (distributed shared memory addressing), HMMA = legacy mma.sync path (must be 0)."
echo
cuobjdump -sass "$SO" 2>/dev/null | awk '
/Function :/ { f=$3; sub(/^_ZN2bg[0-9]*_GLOBAL__N__[0-9a-f]*_[0-9]*_/,"",f); names[f]=1; order[++n]=f }
f!="" {
  if ($0 ~ /UTCHMMA/) { c[f,"UTCHMMA"]++; if ($0 ~ /2CTA/) c[f,"UTCHMMA.2CTA"]++ }
  if ($0 ~ /UTCBAR/) c[f,"UTCBAR"]++
  if ($0 ~ /UTMALDG/) c[f,"UTMALDG"]++
  if ($0 ~ /LDTM/) c[f,"LDTM"]++
  if ($0 ~ /SYNCS/) c[f,"SYNCS"]++
  if ($0 ~ /ACQBULK/) c[f,"ACQBULK"]++
  if ($0 ~ /PREEXIT/) c[f,"PREEXIT"]++
  if ($0 ~ /REDG|RED\.|ATOMG/) c[f,"RED"]++
  if ($0 ~ /LDGSTS/) c[f,"LDGSTS"]++
  if ($0 ~ /UCGABAR/) c[f,"UCGABAR"]++
  if ($0 ~ /MAPA/) c[f,"MAPA"]++
  if ($0 ~ /[^A-Z]HMMA/ && $0 !~ /UTCHMMA/) c[f,"HMMA"]++
}
END {
  printf "%-70s %8s %6s %7s %7s %5s %6s %8s %8s %5s %7s %8s %5s %5s\n","kernel","UTCHMMA","2CTA","UTCBAR","UTMALDG","LDTM","SYNCS","ACQBULK","PREEXIT","RED","LDGSTS","UCGABAR","MAPA","HMMA"
  for (i=1;i<=n;i++) { f=order[i]; g=f; if (length(g)>70) g=substr(g,1,70);
    printf "%-70s %8d %6d %7d %7d %5d %6d %8d %8d %5d %7d %8d %5d %5d\n", g, c[f,"UTCHMMA"], c[f,"UTCHMMA.2CTA"], c[f,"UTCBAR"], c[f,"UTMALDG"], c[f,"LDTM"], c[f,"SYNCS"], c[f,"ACQBULK"], c[f,"PREEXIT"], c[f,"RED"], c[f,"LDGSTS"], c[f,"UCGABAR"], c[f,"MAPA"], c[f,"HMMA"] }
}'
