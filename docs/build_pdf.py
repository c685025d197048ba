"""Renders docs/tutorial.md to docs/tutorial.pdf with nothing but the standard library (the reference ships tuto.pdf; there
is no LaTeX / pandoc in this image).  Text only: headings in Helvetica-Bold, body in Helvetica, code blocks and tables in
Courier; figures are referenced by name (they are SVG files under docs/figs/)."""
import os
import re
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, MARGIN = 612, 792, 54
FONTS = {"body": ("F1", 10.5, 14), "code": ("F3", 8.5, 11), "h1": ("F2", 18, 26), "h2": ("F2", 14, 22), "h3": ("F2", 11.5, 18)}
AVG_W = {"F1": 0.50, "F2": 0.56, "F3": 0.60}      # average glyph width / font size (Courier is exactly 0.6)


def esc(s: str) -> str:
    s = s.encode("latin-1", "replace").decode("latin-1")
    return s.replace("\\", "\\\\").replace("(", "\\(").replace(")", "\\)")


def plain(s: str) -> str:
    s = re.sub(r"!\[([^\]]*)\]\(([^)]+)\)", r"[figure: \2]", s)
    s = re.sub(r"[`*]", "", s)
    for a, b in (("—", "-"), ("–", "-"), ("→", "->"), ("←", "<-"), ("≈", "~"), ("·", "*"), ("−", "-"), ("×", "x"), ("↔", "<->"), ("≤", "<="), ("≥", ">="), ("⇒", "=>"), ("Σ", "sum")):
        s = s.replace(a, b)
    return s


def layout(md: str):
    """-> list of (style, text) lines, already wrapped."""
    out, in_code = [], False
    for raw in md.splitlines():
        if raw.startswith("```"):
            in_code = not in_code
            out.append(("body", ""))
            continue
        if in_code or raw.startswith("|"):
            style, text = "code", plain(raw) if not in_code else raw
        else:
            m = re.match(r"(#+) (.*)", raw)
            if m:
                style, text = ("h1", "h2", "h3")[min(len(m.group(1)), 3) - 1], plain(m.group(2))
                out.append(("body", ""))
            else:
                style, text = "body", plain(raw)
        font, size, _ = FONTS[style]
        width = int((W - 2 * MARGIN) / (size * AVG_W[font]))
        if not text.strip():
            out.append((style, ""))
            continue
        indent = "  " if style == "body" and text.startswith("* ") else ""
        wrapped = textwrap.wrap(text, width=width, subsequent_indent=indent, break_long_words=True, replace_whitespace=False) or [""]
        out += [(style, t) for t in wrapped]
    return out


def paginate(lines):
    pages, cur, y = [], [], H - MARGIN
    for style, text in lines:
        lead = FONTS[style][2]
        if y - lead < MARGIN:
            pages.append(cur)
            cur, y = [], H - MARGIN
        y -= lead
        cur.append((style, text, y))
    if cur:
        pages.append(cur)
    return pages


def build(md: str) -> bytes:
    pages = paginate(layout(md))
    objs = []

    def add(body: bytes) -> int:
        objs.append(body)
        return len(objs)

    catalog = add(b"")                         # 1: filled in below
    pages_obj = add(b"")                       # 2
    f1 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica /Encoding /WinAnsiEncoding >>")
    f2 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica-Bold /Encoding /WinAnsiEncoding >>")
    f3 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Courier /Encoding /WinAnsiEncoding >>")
    kids = []
    for n, page in enumerate(pages):
        ops = ["BT"]
        for style, text, y in page:
            font, size, _ = FONTS[style]
            ops.append(f"/{font} {size} Tf 1 0 0 1 {MARGIN} {y:.1f} Tm ({esc(text)}) Tj")
        ops.append(f"/F1 8 Tf 1 0 0 1 {W // 2 - 10} 30 Tm ({n + 1} / {len(pages)}) Tj")
        ops.append("ET")
        stream = "\n".join(ops).encode("latin-1")
        content = add(b"<< /Length %d >>\nstream\n" % len(stream) + stream + b"\nendstream")
        kids.append(add(b"<< /Type /Page /Parent %d 0 R /MediaBox [0 0 %d %d] /Contents %d 0 R /Resources << /Font << /F1 %d 0 R "
                        b"/F2 %d 0 R /F3 %d 0 R >> >> >>" % (pages_obj, W, H, content, f1, f2, f3)))
    objs[catalog - 1] = b"<< /Type /Catalog /Pages %d 0 R >>" % pages_obj
    objs[pages_obj - 1] = b"<< /Type /Pages /Count %d /Kids [%s] >>" % (len(kids), b" ".join(b"%d 0 R" % k for k in kids))
    out = bytearray(b"%PDF-1.4\n%\xe2\xe3\xcf\xd3\n")
    offsets = []
    for i, body in enumerate(objs, 1):
        offsets.append(len(out))
        out += b"%d 0 obj\n" % i + body + b"\nendobj\n"
    xref = len(out)
    out += b"xref\n0 %d\n" % (len(objs) + 1) + b"0000000000 65535 f \n"
    for off in offsets:
        out += b"%010d 00000 n \n" % off
    out += b"trailer\n<< /Size %d /Root %d 0 R >>\nstartxref\n%d\n%%%%EOF\n" % (len(objs) + 1, catalog, xref)
    return bytes(out)


if __name__ == "__main__":
    md = open(os.path.join(HERE, "tutorial.md")).read()
    pdf = build(md)
    open(os.path.join(HERE, "tutorial.pdf"), "wb").write(pdf)
    print("wrote docs/tutorial.pdf", len(pdf), "bytes")
