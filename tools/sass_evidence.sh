#!/bin/bash
# opcode histogram + the TMA / mbarrier / tensor-memory / dependent-launch / system-scope instructions of one kernel in an object file
#   tools/sass_evidence.sh obj mangled-name-substring title      (needs cuobjdump only, no GPU)
obj=$1; pat=$2; title=$3
echo "== $title"
cuobjdump -sass "$obj" | awk -v pat="$pat" '/Function :/{on=index($0, pat) > 0} on' > /tmp/sass_fn.$$
grep -E '^ +/\*[0-9a-f]{4}\*/' /tmp/sass_fn.$$ | awk '{ i=2; if ($2 ~ /^@/) i=3; print $i }' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END{print ""}'
echo "-- TMA / mbarrier / tensor memory / dependent launch / system-scope instructions:"
grep -E '^ +/\*[0-9a-f]{4}\*/' /tmp/sass_fn.$$ | sed 's/^ *\/\*[0-9a-f]*\*\/ *//; s/ *\/\*.*//; s/;$//' | grep -E 'UBLKCP|SYNCS|LDTM|STTM|UTCATOMSWS|ACQBULK|PREEXIT|MEMBAR|ERRBAR|\.SYS|ATOMG|NANOSLEEP|CCTL' | sed 's/R[0-9]\+/R/g; s/UR[0-9]\+/UR/g; s/UP[0-9]/UP/g; s/P[0-9]/P/g; s/0x[0-9a-f]\+/imm/g' | sort | uniq -c | sort -rn | head -24
rm -f /tmp/sass_fn.$$
