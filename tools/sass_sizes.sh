#!/bin/bash
# static SASS instruction count per kernel of a built library:  bash tools/sass_sizes.sh <lib.so> [name filter]
cuobjdump -sass "$1" 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | grep "${2:-.}" | sort -rn | c++filt -p 2>/dev/null | cut -c1-120
