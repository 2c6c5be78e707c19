"""Turn the output directory of tools/ref_dump (run elsewhere, against the genuine reference crate) into the fixtures
tests/test_reference_golden.py looks for:
    tests/golden/reference_tables/turner2004.tbl, contrafold_v202.tbl     the genuine rna-ss-params blobs
    tests/golden/reference_trna.npz                                        golden vectors of assets/sampled_trnas.fa
usage: python tests/golden/import_reference_dump.py DIR"""
import glob
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(src: str) -> None:
    os.makedirs(os.path.join(HERE, "reference_tables"), exist_ok=True)
    for name in ("turner2004.tbl", "contrafold_v202.tbl"):
        shutil.copy(os.path.join(src, name), os.path.join(HERE, "reference_tables", name))
    arrays = {}
    for path in sorted(glob.glob(os.path.join(src, "*.npy"))):
        arrays[os.path.splitext(os.path.basename(path))[0]] = np.load(path)
    np.savez_compressed(os.path.join(HERE, "reference_trna.npz"), **arrays)
    print(f"imported {len(arrays)} arrays and 2 table blobs")


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    main(sys.argv[1])
