#!/bin/bash
# Installs the UNMODIFIED reference (BindsNET, pure Python) into baseline/_ref (git-ignored; travels to the
# GPU box with gpurun).  The reference declares the poetry build backend, which this image does not have,
# so the install runs from a scratch copy under /tmp whose pyproject.toml names setuptools instead — the
# package sources themselves are not touched.  Dependencies are not resolved (--no-deps): the stack it needs
# on the hot path (torch, numpy) is already in the image; bench.py stub-imports the sub-packages it uses.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refsrc.XXXXXX)"
cp -r "$SRC"/bindsnet "$TMP"/bindsnet
cat > "$TMP"/pyproject.toml <<'EOT'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "bindsnet"
version = "0.3.3"
[tool.setuptools.packages.find]
include = ["bindsnet*"]
EOT
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" 2>&1 | tail -2
rm -rf "$TMP"
ls "$HERE/_ref"
