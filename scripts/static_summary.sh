#!/bin/bash
# Static evidence of the built library (no GPU needed): target arch, per-kernel registers / shared memory / spill stack,
# SASS instruction counts and the mnemonics that show the 64-bit atomic z-test, the programmatic-dependent-launch
# hooks, the 16-byte gathers, the TMA bulk copies + mbarrier waits of the binned variants and the L1 prefetches of the triangle records. Usage: scripts/static_summary.sh > profiles/<round>_static.txt
LIB=${1:-diff-dope_b200/diffdope/_lib/libddope_b200.so}
echo "== $LIB"
cuobjdump --list-elf "$LIB" | sed 's/^/  /'
echo "== resources per kernel (cuobjdump --dump-resource-usage)"
cuobjdump --dump-resource-usage "$LIB" 2>/dev/null | awk '/Function/{f=$2} /REG:/{print "  " f " " $0}' | sed 's/TEXTURE.*//' | c++filt | sed 's/(.*)//'
echo "== SASS instruction counts and selected mnemonics per kernel"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ {if (f!="") printf("  %-60s inst %5d  REDG.MIN.64 %d  LDG.128 %d  STG.128 %d  SHFL %d  VOTE %d  BAR %d  MUFU.RCP %d  PDL(PREEXIT+ACQBULK) %d  TMA(UBLKCP) %d  MBAR(SYNCS) %d  L1-PREFETCH(CCTL.PF) %d\n", f, n, am, l128, s128, sh, vo, ba, rc, pd, tm, mb, pf); f=$3; n=0; am=0; l128=0; s128=0; sh=0; vo=0; ba=0; rc=0; pd=0; tm=0; mb=0; pf=0}
  /^ +\/\*[0-9a-f]+\*\/ / {n++}
  /REDG.*MIN.*64|ATOMG.*MIN.*64/ {am++}
  /LDG\.E\.128|LDG\.E\.CONSTANT\.128|LDG.*\.128/ {l128++}
  /STG.*\.128/ {s128++}
  /SHFL/ {sh++}
  /VOTE/ {vo++}
  /BAR\.SYNC|BAR\.RED/ {ba++}
  /MUFU\.RCP/ {rc++}
  /ACQBULK|PREEXIT/ {pd++}
  /UBLKCP/ {tm++}
  /SYNCS/ {mb++}
  /CCTL.*PF/ {pf++}
  END {printf("  %-60s inst %5d  REDG.MIN.64 %d  LDG.128 %d  STG.128 %d  SHFL %d  VOTE %d  BAR %d  MUFU.RCP %d  PDL(PREEXIT+ACQBULK) %d  TMA(UBLKCP) %d  MBAR(SYNCS) %d  L1-PREFETCH(CCTL.PF) %d\n", f, n, am, l128, s128, sh, vo, ba, rc, pd, tm, mb, pf)}' | c++filt | sed 's/(ddope::SceneDev[^)]*)//'
