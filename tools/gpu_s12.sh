#!/bin/bash
# final evidence + regression of the round-2 code: full GPU suite, smoke, bench, ncu launch list and --set full capture
bash tools/gpu_s9.sh
sed -i 's/s9_/s12a_/g' /dev/null
bash tools/gpu_s6.sh
