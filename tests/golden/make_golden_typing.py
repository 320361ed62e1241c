#!/usr/bin/env python
"""Golden vectors for type_many.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_typing.py

Every assembly of tests/typing_cases.py goes through the UNMODIFIED reference pipeline -- kaptive.serotyping.Serotyper.__call__
(/root/reference/src/kaptive/serotyping/core.py:124-486) and KaptiveRow.from_result (serotyping/io.py:192-296) -- with its `rammappy`
import resolved to the oracle-backed shim (tests/ref_shim).  Stored: the TSV row and the fields of the SerotypingResult."""
import io
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
import ref_bridge as rb  # noqa: E402
import typing_cases as tc  # noqa: E402


def main():
    assert rb.reference_available()
    rb.import_reference("oracle")
    from kaptive.core.genome import GenomeAssembly
    from kaptive.serotyping import Serotyper
    from kaptive.serotyping.io import KaptiveRow

    db, n_k, n_o = tc.make_db()
    st = Serotyper(rb.reference_database(db))
    out = {}
    for a in tc.make_assemblies(db, n_k, n_o):
        contigs = [c for c in a.contigs if len(c[1]) > 0]
        g = GenomeAssembly.from_stream(io.BytesIO(cases.fasta_bytes(contigs)), a.name)
        r = st(g)
        gh = r.gene_hits
        out[a.name] = {
            "row": bytes(KaptiveRow.from_result(r)).decode(),
            "best_locus_idx": int(r.best_locus_idx), "best_locus_score": float(r.best_locus_score), "completeness": float(r.best_locus_completeness),
            "typeable": bool(r.typeable), "problems": int(r.problems), "percent_identity": float(r.percent_identity),
            "percent_coverage": float(r.percent_coverage),
            "length_discrepancy": None if np.isnan(r.length_discrepancy) else float(r.length_discrepancy),
            "gene": gh.gene_indices.tolist(), "t_ctg": gh.t_indices.tolist(), "t_start": gh.t_starts.tolist(), "t_end": gh.t_ends.tolist(),
            "strand": gh.strands.tolist(), "is_expected": gh.is_expected.astype(int).tolist(), "is_inside": gh.is_inside.astype(int).tolist(),
            "is_extra": gh.is_extra.astype(int).tolist(), "state": r.gene_states.tolist(),
            "prot_ident": [float(x) for x in r.protein_identities], "coverage": [float(x) for x in gh.coverages],
            "piece_ctg": r.locus_pieces.ctg_indices.tolist(), "piece_start": r.locus_pieces.starts.tolist(),
            "piece_end": r.locus_pieces.ends.tolist(), "piece_strand": r.locus_pieces.strands.tolist(),
            "missing": list(r.missing_expected_genes), "n_contigs": len(contigs),
        }
        print(a.name, r.best_locus_name, r.typeable, out[a.name]["row"].split("\t")[7], len(gh))
    (HERE / "typing_golden.json").write_text(json.dumps(out, indent=0, sort_keys=True))


if __name__ == "__main__":
    sys.exit(main())
