#!/bin/bash
# regenerate the ICPC specialised kernel (meta device, CPU only) and print its static instruction budget
cd "$(dirname "$0")/.."
CU=$(python - <<'PY'
import yaml, sys
from dspeed_b200 import codegen
cfg = yaml.safe_load(open("dspeed_b200/configs/hpge_icpc.yaml"))
path, text = codegen.prebuild(cfg)
if path is None:
    sys.exit("not specialised: " + text)
print(path.replace(".so", ".cu"))
PY
) || exit 1
echo "$CU"
python scripts/sass_static.py "$CU" "$@"
