#!/bin/bash
# usage: sasscount2.sh file.sass pattern -> per-kernel instruction counts (NOPs excluded)
awk -v pat="$2" '
/Function :/ {f=$3}
/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ { if ($0 ~ /NOP/) next; if (f ~ pat) c[f]++ }
END {for (k in c) print c[k], k}' "$1" | sort -k2
