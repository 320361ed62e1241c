"""The UNMODIFIED reference pipeline (kaptive.serotyping.Serotyper + KaptiveRow, /root/reference/src) on top of THIS repo's mapping
logic.  CPU tier: the product's device code compiled for the host (tests/host_emul) behind a `rammappy`-shaped shim; the TSV rows must
be byte-identical to tests/golden/kaptive_rows.json (made with the oracle behind the same shim).  GPU tier (opt-in, needs the
reference on the box, i.e. the build container with a GPU): the product shim kaptive_b200/shim/rammappy itself.
Skipped where /root/reference does not exist (the GPU box)."""

from __future__ import annotations

import io
import json
import os
from pathlib import Path

import pytest

import cases
import ref_bridge as rb

ROWS = json.loads((Path(__file__).resolve().parent / "golden" / "kaptive_rows.json").read_text())
needs_reference = pytest.mark.skipif(not rb.reference_available(), reason="/root/reference is only present in the build container")


def _run(names, engine):
    os.environ["KB_REF_SHIM_ENGINE"] = engine
    try:
        rb.import_reference("oracle")  # tests/ref_shim: the engine behind it is chosen by KB_REF_SHIM_ENGINE
        from kaptive.core.genome import GenomeAssembly
        from kaptive.serotyping import Serotyper
        from kaptive.serotyping.io import KaptiveRow

        serotypers, out = {}, {}
        for name in names:
            db, contigs = cases.CASES[name]()
            contigs = [c for c in contigs if len(c[1]) > 0]
            if id(db) not in serotypers:
                serotypers[id(db)] = Serotyper(rb.reference_database(db))
            g = GenomeAssembly.from_stream(io.BytesIO(cases.fasta_bytes(contigs)), name)
            res = serotypers[id(db)](g)
            out[name] = (bytes(KaptiveRow.from_result(res)).decode(), res.best_locus_name, bool(res.typeable))
        return out
    finally:
        os.environ.pop("KB_REF_SHIM_ENGINE", None)


@needs_reference
def test_reference_serotyper_over_the_host_compiled_device_code_reproduces_the_golden_rows():
    names = [n for n in cases.CASES if "kaptive_row" in ROWS.get(n, {})]
    assert len(names) >= 15
    got = _run(names, "emul")
    for n in names:
        assert got[n][0] == ROWS[n]["kaptive_row"], n
        assert got[n][1] == ROWS[n]["best_locus"] and got[n][2] == ROWS[n]["typeable"], n


@needs_reference
@pytest.mark.gpu
def test_reference_serotyper_over_the_cuda_shim_reproduces_the_golden_rows():
    """Opt-in by construction: needs both a GPU and /root/reference (never true on the driver's GPU box, true in a dev container
    with a GPU).  Drop-in check of kaptive_b200/shim/rammappy under the reference's own call sites (serotyping/core.py:147-155)."""
    rb.import_reference("gpu")
    from kaptive.core.genome import GenomeAssembly
    from kaptive.serotyping import Serotyper
    from kaptive.serotyping.io import KaptiveRow

    names = [n for n in cases.CASES if "kaptive_row" in ROWS.get(n, {})]
    serotypers = {}
    for name in names:
        db, contigs = cases.CASES[name]()
        contigs = [c for c in contigs if len(c[1]) > 0]
        if id(db) not in serotypers:
            serotypers[id(db)] = Serotyper(rb.reference_database(db))
        g = GenomeAssembly.from_stream(io.BytesIO(cases.fasta_bytes(contigs)), name)
        assert bytes(KaptiveRow.from_result(serotypers[id(db)](g))).decode() == ROWS[name]["kaptive_row"], name
