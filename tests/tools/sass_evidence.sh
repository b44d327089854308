#!/bin/bash
# SASS evidence of the two step kernels (no GPU needed): instruction histograms from the objects of the current build.
# usage: bash tests/tools/sass_evidence.sh > profiles/r1_sass_evidence.md
cd "$(dirname "$0")/../.."
hist() { cuobjdump -sass -fun "$1" "$2" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//' | sed 's/^@!\?U\?P[0-9T] *//' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn; }
SC='_ZN3ion16k_stream_collideILi2ELi2ELb1ELb0ELb0ELb0EEEvNS_5KArgsEfff'
EB='_ZN3ion17k_update_e_b_pairILi16ELi16ELi256ELb1EEEvNS_5KArgsEPKNS_9LodSourceEjNS_10ForeignSetE'
echo "# r1 — SASS of the two step kernels (\`cuobjdump -sass\` on build/csrc/*.o, sm_100a; \`bash tests/tools/sass_evidence.sh\`)"
echo
echo "What to look for: packed FP32 (\`FFMA2\` / \`FADD2\` / \`FMUL2\`), loads pinned in front of the flag branch (\`LDG.E.STRONG.SYS\` = \`ld.volatile.global\`),"
echo "the 16-byte vector reduction of the LOD deposit (\`REDG.E.ADD.F32x4\`), and NO tensor instructions (\`HMMA\`, \`UTC*MMA\`): the path is HBM- / FP32-issue bound."
for spec in "k_stream_collide<D3Q19, FP32, MHD, SRT, even step>:$SC:build/csrc/sc_d3q19.o" "k_update_e_b_pair<16, 16, 256>:$EB:build/csrc/fields.o"; do
  name=${spec%%:*}; rest=${spec#*:}; fun=${rest%%:*}; obj=${rest#*:}
  total=$(hist "$fun" "$obj" | awk '{s+=$1} END {print s}')
  echo; echo "## \`$name\` — $total instructions"; echo; echo '```'
  hist "$fun" "$obj" | awk '$1>=8 {printf "%6d %s\n", $1, $2}'
  echo '```'
  echo; echo "tensor instructions: $(cuobjdump -sass -fun "$fun" "$obj" | grep -cE 'HMMA|UTC[A-Z]*MMA|HGMMA')"
done
echo; echo "Registers / spills (\`-Xptxas -v\`): stream_collide MHD FP32 127 registers, 0 spill bytes; update_e_b_pair 255 registers (1 block of 256 threads per SM by design, 128 KB shared memory)."
