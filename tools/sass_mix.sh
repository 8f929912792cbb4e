#!/bin/bash
# usage: tools/sass_mix.sh file.cubin|.so [function-substring]  -> per-function SASS opcode histogram (top 16)
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ {name=$3}
/^ *\/\*[0-9a-f]+\*\/ +[A-Z@!]/ {
  if (pat != "" && index(name, pat) == 0) next;
  op=$2; if (op ~ /^@/) op=$3; split(op,a,";"); op=a[1];
  n=split(op,b,"."); key=b[1]; if (b[2]=="WIDE"||b[2]=="HI"||b[2]=="X"||b[2]=="MOV") key=key"."b[2]; if (b[3]=="X"||b[4]=="X") key=key".X";
  cnt[name" "key]++; tot[name]++ }
END { for (k in cnt) print k, cnt[k]; for (n in tot) print n, "TOTAL", tot[n] }' | sort -k1,1 -k3,3nr | awk '{c[$1]++; if (c[$1]<=16) print}'
