"""TEST / BASELINE INFRASTRUCTURE ONLY — stage the unmodified Python reference so it can travel to the GPU box.

`/root/reference` exists only in the build container.  The reference is pure Python (nothing to compile), so the
"build" of oracle/_ref is an archive: this recipe zips the reference's Python packages, configs and tokenizer
vocabulary as they lie under /root/reference into `oracle/_ref/reference.zip` (git-ignored: reference sources never
enter the history; not gpurun-ignored: it ships with the snapshot like a built .so).  `oracle/shim.py` unpacks it into
a temporary directory when /root/reference is absent, so `bench.py --impl reference` and the `cpu_baseline` /
`reference_gpu_eager` legs time the REAL reference modules (kind "reference") instead of the restated port.

    python oracle/stage_ref.py            # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import sys
import zipfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref" / "reference.zip"
KEEP_TOP = ("Dassl", "clip", "configs", "datasets", "evaluation", "trainers", "utils", "federated_main.py",
            "requirements.txt")
SKIP_SUFFIX = (".pyc", ".png", ".jpg", ".gif", ".pdf")


def stage(ref_root: str = "/root/reference", out: Path = OUT) -> Path | None:
    root = Path(ref_root)
    if not (root / "trainers").is_dir():
        return None
    files = []
    for top in KEEP_TOP:
        p = root / top
        if p.is_file():
            files.append(p)
        elif p.is_dir():
            files += [f for f in sorted(p.rglob("*")) if f.is_file() and "__pycache__" not in f.parts
                      and not f.name.endswith(SKIP_SUFFIX)]
    newest = max(f.stat().st_mtime for f in files)
    if out.exists() and out.stat().st_mtime >= newest:
        return out
    out.parent.mkdir(parents=True, exist_ok=True)
    tmp = out.with_suffix(".tmp")
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, f.relative_to(root).as_posix())
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    res = stage(*sys.argv[1:2])
    print(res if res else "reference tree not found: nothing staged")
