"""Test support: restatements of `bean run` pre-/post-processing that is OUT OF SCOPE of the hot path (SURVEY section 2, rows 4, 7,
8, 14) but needed by the parity harness -- screen preparation (`prepare_bdata`), the CLI's argument checks and per-target tables,
the tiling result-table summaries, and a pure-Python bigWig reader for `--acc-bw-path` (pyBigWig is not installable here).
They were written in round 1 inside the package; they live here because the product is the SVI path, not the CLI around it."""
