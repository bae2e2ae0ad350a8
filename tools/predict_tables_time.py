"""Tables in -> tables out on a synthetic genes / features table pair (tools/tables_time.py's generator): wall time of
gecco_b200.tables.predict_tables (load, pack, marginals on the B200, genes / features / clusters tables) with the
feature extraction on the host and on the device.  B200 only.

    python tools/predict_tables_time.py [genes] [rows_per_gene]
"""
import os
import pathlib
import sys
import tempfile
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from gecco_b200 import model_io
from gecco_b200.crf import ClusterCRF
from gecco_b200.tables import predict_tables
from tools.tables_time import write_tables

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
rows = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
w = model_io.load_tsv_model(model_io.bundled_model_dir())
tmp = pathlib.Path(tempfile.mkdtemp(prefix="gcrf_predict_"))
t0 = time.perf_counter()
gpath, fpath = write_tables(tmp, genes, rows, max(1, genes // 40), w.attrs)
print(f"generated {genes} genes in {time.perf_counter() - t0:.1f} s", flush=True)
crf = ClusterCRF.trained(None)
ref = None
for mode in ("1", "0", "1", "0"):
    os.environ["GECCO_B200_HOST_FEATURES"] = mode
    t0 = time.perf_counter()
    tables, prob = predict_tables(gpath, fpath, tmp / f"out{mode}", model=crf)
    dt = time.perf_counter() - t0
    what = "host packer" if mode == "1" else "features on device"
    print(f"{what}: {tables.genes} genes / {tables.domains} rows in {dt:.3f} s = {tables.genes / dt / 1e6:.2f} M genes/s", flush=True)
    tables.close()
same = all((tmp / "out0" / n).read_bytes() == (tmp / "out1" / n).read_bytes() for n in os.listdir(tmp / "out0"))
print("result tables identical (host vs device feature extraction):", same, sorted(os.listdir(tmp / "out0")))
