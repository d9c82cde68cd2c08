"""FASTA-only stand-in for Bio.SeqIO (parse / write) - test infrastructure, see ../README.md."""
_COMP = str.maketrans("ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")


class Seq:
    def __init__(self, data):
        self._d = str(data)

    def __str__(self):
        return self._d

    def __len__(self):
        return len(self._d)

    def __getitem__(self, idx):
        r = self._d[idx]
        return Seq(r) if isinstance(idx, slice) else r

    def upper(self):
        return Seq(self._d.upper())

    def reverse_complement(self):
        return Seq(self._d.translate(_COMP)[::-1])


class SeqRecord:
    def __init__(self, seq, id="", description=""):
        self.seq, self.id, self.name, self.description = seq, id, id, description

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, idx):
        return SeqRecord(self.seq[idx], self.id, self.description)


def parse(handle, fmt="fasta"):
    assert fmt == "fasta"
    close = False
    if not hasattr(handle, "read"):
        handle, close = open(handle, "r"), True
    try:
        header, chunks = None, []
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if header is not None:
                    yield SeqRecord(Seq("".join(chunks)), header.split()[0] if header.split() else "", header)
                header, chunks = line[1:], []
            elif header is not None:
                chunks.append(line.strip())
        if header is not None:
            yield SeqRecord(Seq("".join(chunks)), header.split()[0] if header.split() else "", header)
    finally:
        if close:
            handle.close()


def write(records, handle, fmt="fasta"):
    assert fmt == "fasta"
    close = False
    if not hasattr(handle, "write"):
        handle, close = open(handle, "w"), True
    n = 0
    for rec in records:
        s = str(rec.seq)
        handle.write(">%s\n" % (rec.description or rec.id))
        for i in range(0, len(s), 60):
            handle.write(s[i : i + 60] + "\n")
        n += 1
    if close:
        handle.close()
    return n
