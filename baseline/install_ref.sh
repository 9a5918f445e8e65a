#!/bin/bash
# Put the UNMODIFIED reference (Wuziyi616/SlotDiffusion) under baseline/_ref so that it travels to the GPU box
# (baseline/_ref is git-ignored, NOT gpurun-ignored; /root/reference does not exist on the box).
#
#   bash baseline/install_ref.sh            # in the build container, where /root/reference exists
#
# Step 1 is the prescribed offline install.  It "succeeds" but produces an EMPTY wheel: the reference's top-level
# `slotdiffusion/` directory has no __init__.py (it is an implicit namespace package), so setup.py's find_packages()
# returns [] and only slotdiffusion-0.1.0.dist-info is installed.  Step 2 therefore places the package tree itself
# (python files only -- no dataset split lists / pickles) next to that dist-info, byte for byte, which is what the wheel
# would have contained had find_packages() seen the namespace package.  Nothing from baseline/_ref is ever committed.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${SDB_REFERENCE_ROOT:-/root/reference}
[ -d "$REF/slotdiffusion" ] || { echo "reference not found at $REF"; exit 1; }
rm -rf "$HERE/_ref" /tmp/_sdb_refcopy
cp -r "$REF" /tmp/_sdb_refcopy            # /root/reference is read-only; the build writes egg-info into the tree
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" /tmp/_sdb_refcopy 2>&1 | tail -2
(cd "$REF" && find slotdiffusion -name '*.py' | while read -r f; do mkdir -p "$HERE/_ref/$(dirname "$f")"; cp "$f" "$HERE/_ref/$f"; done)
rm -rf /tmp/_sdb_refcopy
# integrity: every python file identical to the reference's
(cd "$REF" && find slotdiffusion -name '*.py' | while read -r f; do cmp -s "$f" "$HERE/_ref/$f" || { echo "MISMATCH $f"; exit 1; }; done)
echo "baseline/_ref: $(find "$HERE/_ref/slotdiffusion" -name '*.py' | wc -l) python files, $(du -sh "$HERE/_ref" | cut -f1)"
