#!/bin/bash
# cuobjdump -sass of the two pair kernels as shipped (the instances the bench runs) -> profiles/<tag>_sass_k_*.txt
# usage: scripts/sass_listing.sh [tag]      (no GPU needed)
cd "$(dirname "$0")/.."
TAG=${1:-r02}
SO=pi_sph_fluid_b200/libsphb200.so
for k in k_density k_force; do
  pat="k_densityILb0ELb0ELb1E"; [ $k = k_force ] && pat="k_forceILb0ELb1ELb1ELb0ELi1E"
  fn=$(cuobjdump -sass $SO 2>/dev/null | grep -o "Function : .*$pat.*" | head -1 | sed 's/Function : //')
  body=$(cuobjdump -sass -fun "$fn" $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/")
  { echo "# cuobjdump -sass of the shipped libsphb200.so, $k, the instance the bench runs:"; echo "# $fn"
    echo "# mnemonic counts (UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier, FADD2/FMUL2/FFMA2 = packed fp32, PREEXIT = PDL):"
    echo "$body" | sed 's#/\*[0-9a-f]*\*/##g; s#/\* 0x[0-9a-f]* \*/##' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{printf "#   %6d %s\n", $1, $2}'
    echo "$body" | sed 's#/\* 0x[0-9a-f]* \*/##; s/ *$//'; } > profiles/${TAG}_sass_$k.txt
  wc -l profiles/${TAG}_sass_$k.txt
done
