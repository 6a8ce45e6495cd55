#!/bin/bash
# usage: sasscount.sh file.cubin  -> instruction count per kernel (and per-opcode histogram with -v)
cuobjdump -sass "$1" | awk -v verbose="$2" '
/Function :/ {f=$3}
/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ {
  if ($0 ~ /NOP/) next; c[f]++;
  op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); h[f" "op]++
}
END {for (k in c) print c[k], k; if (verbose) for (k in h) print "   ", h[k], k}' | sort -k2
