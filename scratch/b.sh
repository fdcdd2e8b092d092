#!/bin/bash
# build libgnnfp.so from the repo root; fail loudly
cd /root/repo || exit 1
python -c "
from gnnkeras_b200 import build
build.build(force=True)" 2>&1 | grep -iE "error|No module" && exit 1
exit 0
