#!/bin/bash
# total SASS instruction count + mnemonic histogram of one kernel in an object file (static proxy for the per-line loop cost)
#   tools/sass_count.sh octproz_b200/csrc/k_fused_r1.o 'ILi1ELi3ELb0ELi0ELb0E' [top]
obj=$1; pat=$2; top=${3:-14}
cuobjdump -sass "$obj" | awk -v pat="$pat" '/Function :/{on=index($0, pat) > 0} on' | grep -E '^ +/\*[0-9a-f]{4}\*/' | awk '{print $2}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn > /tmp/sass_hist.$$
echo "total $(awk '{s+=$1} END{print s}' /tmp/sass_hist.$$)"; head -$top /tmp/sass_hist.$$ | tr '\n' ' '; echo; rm -f /tmp/sass_hist.$$
