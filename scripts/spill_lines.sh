#!/bin/bash
# usage: scripts/spill_lines.sh <object.o> [function-substring]
# local-memory (spill) instructions per source line (-lineinfo builds), per function
set -e
d=$(mktemp -d)
( cd "$d" && cuobjdump -xelf all "$1" >/dev/null && nvdisasm -g -c ./*.cubin > dis.sass 2>/dev/null )
python3 - "$d/dis.sass" "${2:-}" <<'PY'
import re, sys
cur, fn, cnt = None, None, {}
for l in open(sys.argv[1]):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        fn = m.group(1); cur = None; continue
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.search(r'\b(STL|LDL)', l):
        cnt.setdefault(fn, {}).setdefault(cur, 0)
        cnt[fn][cur] += 1
for f, c in cnt.items():
    if sys.argv[2] and sys.argv[2] not in f:
        continue
    print(f[-60:], {k: c[k] for k in sorted(c, key=lambda x: (x is None, x))})
PY
rm -rf "$d"
