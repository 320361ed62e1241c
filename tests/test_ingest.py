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


# ---------------------------------------------------------------------------------------------- packed ingest
_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _CODE[_c | 0x20] = _i
_CODE[ord("U")] = _CODE[ord("u")] = 3


def unpack(pb, c):
    """Contig c of a PackedBatch back to nt4 codes (4 = ambiguous)."""
    so, ln = int(pb.contig_soff[c]), int(pb.contig_len[c])
    b = so + np.arange(ln, dtype=np.int64)
    code = (pb.seq2[b >> 4] >> (2 * (b & 15)).astype(np.uint32)) & 3
    amb = (pb.nmask[b >> 5] >> (b & 31).astype(np.uint32)) & 1
    return np.where(amb == 1, 4, code).astype(np.uint8)


def make_files_with_ambiguity():
    rng = np.random.default_rng(5)
    files = make_files()
    parts = []
    for c in range(6):
        s = np.frombuffer(synth.random_dna(rng, int(rng.integers(1, 3000))).tobytes(), np.uint8).copy()
        s[rng.integers(0, len(s), size=max(1, len(s) // 50))] = np.frombuffer(b"NnRYKMSWUu-*", np.uint8)[rng.integers(0, 12, size=max(1, len(s) // 50))]
        w = int(rng.choice([60, 61, 80, 127, 128, 129]))
        parts.append(b">amb%d\n" % c)
        parts.extend(s[k : k + w].tobytes() + b"\n" for k in range(0, len(s), w))
    files.insert(3, b"".join(parts))
    return files


@pytest.mark.parametrize("threads,simd", [(1, True), (7, True), (3, False)])
def test_packed_ingest_equals_the_ascii_path(threads, simd):
    """kb_fasta_ingest_count / _lengths / kb_packed_layout / kb_fasta_ingest_pack: every contig, unpacked, equals the nt4 codes of
    the bytes the ASCII reader yields; padding is ambiguous, sequence words of padding are zero; the AVX2 + BMI2 packer and the
    portable loop are bit-identical; the layout is the device's (contigs on 128-base boundaries, 128 padded bases at both ends)."""
    files = make_files_with_ambiguity()
    ref = ingest.ingest_fasta(files, threads=2)
    pb = ingest.ingest_fasta_packed(files, threads=threads, use_simd=simd)
    assert np.array_equal(pb.asm_contig_start, ref.asm_contig_start) and np.array_equal(pb.contig_len, ref.contig_len)
    assert pb.names == ref.names
    assert pb.storage_bases % 128 == 0 and len(pb.seq2) == pb.storage_bases // 16 and len(pb.nmask) == pb.storage_bases // 32
    covered = np.zeros(pb.storage_bases, bool)
    for c in range(len(pb.contig_len)):
        so, ln = int(pb.contig_soff[c]), int(pb.contig_len[c])
        assert so % 128 == 0 and so >= 128
        o = int(ref.contig_off[c])
        assert np.array_equal(unpack(pb, c), _CODE[ref.seqs[o : o + ln]]), c
        covered[so : so + ln] = True
    # everything that is not a base of some contig: mask bit set, sequence bits zero
    b = np.nonzero(~covered)[0]
    assert np.all((pb.nmask[b >> 5] >> (b & 31).astype(np.uint32)) & 1 == 1)
    assert np.all((pb.seq2[b >> 4] >> (2 * (b & 15)).astype(np.uint32)) & 3 == 0)
    other = ingest.ingest_fasta_packed(files, threads=2, use_simd=not simd)
    assert np.array_equal(other.seq2, pb.seq2) and np.array_equal(other.nmask, pb.nmask)


def test_packed_ingest_into_caller_buffers_and_empty_input():
    files = make_files()
    pb0 = ingest.ingest_fasta_packed(files, threads=2)
    seq2, nmask = np.full(len(pb0.seq2) + 100, 0xDEADBEEF, np.uint32), np.full(len(pb0.nmask) + 100, 0x12345678, np.uint32)
    pb = ingest.ingest_fasta_packed(files, threads=3, out=(seq2, nmask))
    assert np.array_equal(pb.seq2, pb0.seq2) and np.array_equal(pb.nmask, pb0.nmask)
    assert seq2[len(pb0.seq2)] == 0xDEADBEEF and nmask[len(pb0.nmask)] == 0x12345678  # nothing written past the layout
    e = ingest.ingest_fasta_packed([], threads=1)
    assert e.storage_bases == 256 and np.all(e.nmask == 0xFFFFFFFF) and np.all(e.seq2 == 0)
    e = ingest.ingest_fasta_packed([b"", b">x\n", b""], threads=2)
    assert e.asm_contig_start.tolist() == [0, 0, 1, 1] and e.contig_len.tolist() == [0] and np.all(e.nmask == 0xFFFFFFFF)


def test_read_fasta_files_follows_the_reference_openers(tmp_path):
    """File-name rules and compression formats of GenomeAssembly.from_file (reference core/genome.py:105-106,194-214)."""
    import bz2
    import gzip
    import lzma

    data = b">c1 d\nACGTNNAC\nGT\n>c2\nTTTT\n"
    (tmp_path / "a.fasta").write_bytes(data)
    (tmp_path / "b.fna.gz").write_bytes(gzip.compress(data))
    (tmp_path / "c.fa.bz2").write_bytes(bz2.compress(data))
    (tmp_path / "d.fas.xz").write_bytes(lzma.compress(data))
    paths = [tmp_path / n for n in ("a.fasta", "b.fna.gz", "c.fa.bz2", "d.fas.xz")]
    blobs, ids = ingest.read_fasta_files(paths, threads=3)
    assert blobs == [data] * 4 and ids == ["a", "b", "c", "d"]
    pb = ingest.ingest_fasta_packed(blobs, threads=2)
    assert pb.contig_len.tolist() == [10, 4] * 4 and pb.names[1] == ["c1", "c2"]
    assert unpack(pb, 0).tolist() == [0, 1, 2, 3, 4, 4, 0, 1, 2, 3]
    (tmp_path / "e.txt").write_bytes(data)
    with pytest.raises(NotImplementedError):
        ingest.read_fasta_files([tmp_path / "e.txt"])
