"""Reference arm of the benchmark: runs the UNMODIFIED fawnliu/TRIS sources (never this repo's kernels).

`install_reference.py` copies the reference's Python sources from /root/reference into the git-ignored
`baseline/_ref/` so that they travel to the GPU box with the gpurun snapshot (the reference is a script directory with
no setup.py / pyproject.toml, so `pip install` does not apply).  Nothing under `tris_b200/` imports this package.
"""
