"""CPU tier: batch FASTA ingest (kb_fasta_ingest_count / kb_fasta_ingest_parse) against the single-file reader and a
plain-Python statement of the record semantics (header up to the first whitespace, lines joined, CR / trailing blanks dropped)."""

import numpy as np
import pytest

from kaptive_b200 import ingest, synth


def py_parse(data: bytes):
    recs = []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            name = line[1:].split(b" ")[0].split(b"\t")[0].split(b"\r")[0]
            recs.append([name.decode(), b""])
        elif recs:
            recs[-1][1] += line.rstrip(b"\r \t")
    return [(n, s) for n, s in recs]


def make_files():
    rng = np.random.default_rng(3)
    files = []
    for i in range(23):
        parts = []
        for c in range(int(rng.integers(0, 9))):
            s = synth.random_dna(rng, int(rng.integers(0, 5000))).tobytes()
            if c % 3 == 0:
                s = s.lower()
            w = int(rng.choice([60, 80, 10_000]))
            eol = b"\r\n" if (i + c) % 4 == 0 else b"\n"
            parts.append(b">ctg%d_%d some description\tmore" % (i, c) + eol)
            parts.extend(s[k : k + w] + eol for k in range(0, len(s), w))
            if c % 5 == 1:
                parts.append(eol)  # blank line inside a record
        data = b"".join(parts)
        if i % 2 == 0 and data.endswith(b"\n"):
            data = data[:-1]  # no trailing newline
        files.append(data)
    files.append(b"")                      # empty file
    files.append(b"no header line\nACGT\n")  # sequence before any header is ignored
    return files


@pytest.mark.parametrize("threads", [1, 4, 16])
def test_batch_ingest_matches_reference_semantics(threads):
    files = make_files()
    b = ingest.ingest_fasta(files, threads=threads)
    assert len(b.asm_contig_start) == len(files) + 1 and b.asm_contig_start[0] == 0
    for i, f in enumerate(files):
        want = py_parse(f)
        c0, c1 = int(b.asm_contig_start[i]), int(b.asm_contig_start[i + 1])
        assert c1 - c0 == len(want)
        assert b.names[i] == [n for n, _ in want]
        for k, (_, s) in enumerate(want):
            o, ln = int(b.contig_off[c0 + k]), int(b.contig_len[c0 + k])
            assert ln == len(s) and b.seqs[o : o + ln].tobytes() == s
    assert int(b.contig_len.sum()) == len(b.seqs)


def test_batch_ingest_equals_single_file_reader():
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "kaptive_b200" / "shim"))
    from rammappy.fasta import parse_fasta_bytes

    files = make_files()
    b = ingest.ingest_fasta(files, threads=3)
    for i, f in enumerate(files):
        recs = parse_fasta_bytes(f)
        c0 = int(b.asm_contig_start[i])
        assert [n for n, _ in recs] == b.names[i]
        for k, (_, s) in enumerate(recs):
            o, ln = int(b.contig_off[c0 + k]), int(b.contig_len[c0 + k])
            assert b.seqs[o : o + ln].tobytes() == s


def test_preallocated_output_buffer_and_empty_batch():
    files = make_files()[:5]
    total = sum(len(s) for f in files for _, s in py_parse(f))
    buf = np.zeros(total + 100, np.uint8)
    b = ingest.ingest_fasta(files, threads=2, out=buf)
    assert b.seqs.base is buf or b.seqs.ctypes.data == buf.ctypes.data
    with pytest.raises(ValueError):
        ingest.ingest_fasta(files, out=np.zeros(3, np.uint8))
    e = ingest.ingest_fasta([])
    assert len(e.seqs) == 0 and list(e.asm_contig_start) == [0]
