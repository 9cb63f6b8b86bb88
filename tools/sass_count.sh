#!/bin/bash
# instruction count per kernel of a built library (static code size; pair with -Xptxas -v for registers/spills)
cuobjdump -sass "${1:-deeppowers_b200/libdpfhe.so}" 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -k2 | cut -c1-110
