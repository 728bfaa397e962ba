"""Doc hygiene: every repository path that README.md / DESIGN.md / INTEGRATION.md / profiles/README.md cite in
backticks exists (reference paths and built artefacts are exempt)."""
from __future__ import annotations

import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIXES = ("tools/", "profiles/", "tests/", "scpn_fusion_core_b200/", "oracle/", "include/", "csrc/")
BUILT = (".so", ".sha256", "_ref", "_ref/")
REFERENCE_SIDE = ("tools/parallel_gen_iter.py",)  # cited as the reference's tool, same prefix as this repo's tools/


def test_cited_repo_paths_exist():
    missing = []
    for doc in ("README.md", "DESIGN.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc), encoding="utf-8").read()
        for tok in re.findall(r"`([^`\s]+)`", text):
            tok = re.sub(r":\d[\d,\-]*$", "", tok.split("::")[0].rstrip(".,;:"))
            if not tok.startswith(PREFIXES) or any(tok.endswith(b) for b in BUILT) or "*" in tok or "<" in tok or tok in REFERENCE_SIDE:
                continue
            cands = [tok, os.path.join("scpn_fusion_core_b200", tok), os.path.join("profiles", tok)]
            if not any(os.path.exists(os.path.join(ROOT, c)) for c in cands):
                missing.append((doc, tok))
    assert not missing, missing
