def parse_fasta_bytes(data: bytes):
    """Plain-Python restatement used only for golden generation."""
    out, name, chunks = [], None, []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(chunks)))
            name = line[1:].split()[0].decode("ascii", "replace") if len(line) > 1 and line[1:].split() else ""
            chunks = []
        elif name is not None:
            chunks.append(line.rstrip(b"\r \t"))
    if name is not None:
        out.append((name, b"".join(chunks)))
    return out
