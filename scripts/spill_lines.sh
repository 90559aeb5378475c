#!/bin/bash
# usage: scripts/spill_lines.sh <object.o>  — local-memory (spill) instructions per source line (-lineinfo builds)
set -e
d=$(mktemp -d)
( cd "$d" && cuobjdump -xelf all "$1" >/dev/null && nvdisasm -g -c ./*.cubin > dis.sass 2>/dev/null )
python3 - "$d/dis.sass" <<'PY'
import re, sys
cur, cnt = None, {}
for l in open(sys.argv[1]):
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.search(r'\b(STL|LDL)', l):
        cnt[cur] = cnt.get(cur, 0) + 1
print({k: cnt[k] for k in sorted(cnt)})
PY
rm -rf "$d"
