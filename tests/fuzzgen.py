"""Seeded adversarial FASTA / FASTQ generators shared by the emulation (CPU) and GPU parity tests."""
import random

DNA = b"ACGTacgtNnRY-*>@+ \r"
DNA_W = [30, 30, 30, 30, 6, 6, 6, 6, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]
PROT = b"ACDEFGHIKLMNOPQRSTUVWYacdxBJZX*-> \r"


def rand_seq(rng, n, protein):
    if protein:
        return bytes(rng.choice(PROT) for _ in range(n))
    return bytes(rng.choices(DNA, weights=DNA_W, k=n))


def rand_name(rng):
    n = rng.choice([0, 0, 1, 2, 5, 12, 40])
    alphabet = b"abcXYZ0189 _|>@+\t\r" + (b'"' if rng.random() < 0.15 else b"")
    s = bytes(rng.choice(alphabet) for _ in range(n))
    if rng.random() < 0.1 and n:
        s = b'"' + s + b'"'
    return s


def fasta(rng, protein=False, max_records=8, max_len=300):
    out = bytearray()
    if rng.random() < 0.3:
        out += rand_seq(rng, rng.randrange(0, 30), protein).replace(b">", b"A") + b"\n"  # junk before the first header
    for _ in range(rng.randrange(0, max_records + 1)):
        out += b">" + rand_name(rng).replace(b"\n", b"") + b"\n"
        total = rng.choice([0, 1, 3, 10, 50, max_len])
        total = rng.randrange(0, total + 1)
        width = rng.choice([1, 2, 7, 60, 70, 10 ** 6])
        pos = 0
        while pos < total:
            w = min(width, total - pos)
            line = rand_seq(rng, w, protein)
            if line.startswith(b">"):
                line = b"A" + line[1:]
            out += line
            pos += w
            out += rng.choice([b"\n", b"\n", b"\n", b"\r\n", b"\n\n", b"\n\n\n"])
    if out and rng.random() < 0.3:
        while out and out[-1:] in b"\r\n":
            del out[-1:]
    return bytes(out)


def fastq(rng, protein=False, max_records=10, max_len=200, malformed=0.0):
    out = bytearray()
    for i in range(rng.randrange(0, max_records + 1)):
        n = rng.randrange(0, rng.choice([1, 5, 20, 150, max_len]) + 1)
        seq = rand_seq(rng, n, protein).replace(b"\n", b"")
        name = rand_name(rng).replace(b"\n", b"")
        qual = bytes(rng.choice(b"@+I5#") for _ in range(rng.choice([n, n, 0, 3])))
        tagc = b"@" if rng.random() >= malformed else rng.choice([b"", b"x", b">"])
        plusc = b"+" if rng.random() >= malformed else rng.choice([b"", b"-"])
        eol = rng.choice([b"\n", b"\n", b"\n", b"\r\n"])
        out += tagc + name + eol + seq + eol + plusc + rng.choice([b"", name]) + eol + qual + eol
    r = rng.random()
    if out and r < 0.15:
        out = out[:-1]  # no final line feed
    elif out and r < 0.45:
        cut = rng.randrange(0, len(out))
        out = out[: len(out) - min(cut, rng.choice([1, 2, 5, 30, 400]))]  # truncated file
    return bytes(out)
