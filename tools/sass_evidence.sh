#!/bin/bash
# Regenerate profiles/r1_sass_evidence.txt from the shipped library (no GPU needed).
set -e
cd "$(dirname "$0")/.."
LIB=lshrs_b200/_lib/liblshx.so
OUT=${1:-profiles/r1_sass_evidence.txt}
{
  echo "# cuobjdump -sass $LIB | mnemonic counts (proof of tcgen05 / TMEM / TMA in the shipped library)"
  echo "# nvcc $(nvcc --version | grep -o 'Build.*')"
  cuobjdump -sass $LIB | grep -oE '\b(FFMA|UTCHMMA(\.2CTA)?|UTCQMMA|UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|LDTM|STTM|UTCBAR[.A-Z0-9_]*|SYNCS\.[A-Z0-9_.]+|ELECT|UTCATOMSWS[.A-Z_]*|UCGABAR_[A-Z]+|F2FP[.A-Z0-9_]*|LDG\.E[.A-Z0-9_]*|STG\.E[.A-Z0-9_]*)\b' \
    | sort | uniq -c | sort -rn
  echo
  echo "# kernels"
  cuobjdump -sass $LIB | grep -E "^\s*Function :" | sed 's/^\s*//'
  echo
  echo "# ptxas -v (registers / shared memory)"
  python -m lshrs_b200._build --force -v 2>&1 | awk '/Compiling entry function/ {name=$6} /Used [0-9]+ registers/ {sub(/^ptxas info *: */, ""); print name "\t" $0}' | tr -d "'"
} > $OUT
echo wrote $OUT
