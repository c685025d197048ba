"""Draws the tutorial's diagrams as SVG (counterpart of the reference's `figs/` directory: send_recv, broadcast,
scatter, gather, reduce, all_reduce, all_gather -- tuto.md:84-86,130-142 show them as tables of images).

The pictures are generated, not stored: `python docs/make_figs.py` writes docs/figs/*.svg.  Besides the seven
classic collectives there are three B200-specific drawings (ring all-reduce, one-shot / two-shot over peer memory,
NVLS in-switch reduction) and the layout of one fused training step.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "figs")

COL = ["#4e79a7", "#f28e2b", "#59a14f", "#e15759"]
GREY = "#d0d0d0"


class Svg:
    def __init__(self, w, h, title):
        self.w, self.h = w, h
        self.prims = []          # the same drawing as plain primitives (PNG preview, see `preview_png`)
        self.parts = [
            f'<svg xmlns="http://www.w3.org/2000/svg" width="{w}" height="{h}" viewBox="0 0 {w} {h}" '
            f'font-family="Helvetica,Arial,sans-serif" font-size="13">',
            f"<title>{title}</title>",
            '<defs><marker id="arr" viewBox="0 0 10 10" refX="9" refY="5" markerWidth="7" markerHeight="7" '
            'orient="auto-start-reverse"><path d="M0,0 L10,5 L0,10 z" fill="#333"/></marker></defs>',
            f'<rect width="{w}" height="{h}" fill="white"/>',
        ]

    def box(self, x, y, w, h, fill, label="", stroke="#333", fg="white"):
        self.parts.append(f'<rect x="{x}" y="{y}" width="{w}" height="{h}" rx="4" fill="{fill}" stroke="{stroke}"/>')
        self.prims.append(("rect", x, y, w, h, fill, stroke))
        if label:
            self.text(x + w / 2, y + h / 2 + 4, label, fg, "middle")

    def text(self, x, y, s, fill="#222", anchor="start", size=None, bold=False):
        extra = (f' font-size="{size}"' if size else "") + (' font-weight="bold"' if bold else "")
        self.parts.append(f'<text x="{x}" y="{y}" fill="{fill}" text-anchor="{anchor}"{extra}>{s}</text>')
        self.prims.append(("text", x, y, s, fill, anchor, size or 13))

    def arrow(self, x1, y1, x2, y2, dash=False, color="#333"):
        d = ' stroke-dasharray="5,4"' if dash else ""
        self.parts.append(
            f'<line x1="{x1}" y1="{y1}" x2="{x2}" y2="{y2}" stroke="{color}" stroke-width="1.6"{d} marker-end="url(#arr)"/>')
        self.prims.append(("line", x1, y1, x2, y2, color))

    def save(self, name):
        os.makedirs(OUT, exist_ok=True)
        path = os.path.join(OUT, name + ".svg")
        with open(path, "w") as f:
            f.write("\n".join(self.parts) + "\n</svg>\n")
        if os.environ.get("FIGS_PNG_PREVIEW"):
            preview_png(self, os.path.join(os.environ["FIGS_PNG_PREVIEW"], name + ".png"))
        return path


def preview_png(svg: "Svg", path: str, scale: int = 2) -> None:
    """Rasterise the drawing with PIL (approximate fonts) -- a way to LOOK at the figures where no SVG viewer exists."""
    from PIL import Image, ImageDraw, ImageFont
    img = Image.new("RGB", (svg.w * scale, svg.h * scale), "white")
    d = ImageDraw.Draw(img)

    def font(sz):
        for cand in (os.environ.get("FIGS_PREVIEW_FONT"), "DejaVuSans.ttf"):
            try:
                if cand:
                    return ImageFont.truetype(cand, int(sz * scale))
            except OSError:
                pass
        return ImageFont.load_default()

    import html as _html
    for pr in svg.prims:
        if pr[0] == "rect":
            _, x, y, w, h, fill, stroke = pr
            d.rectangle([x * scale, y * scale, (x + w) * scale, (y + h) * scale], fill=fill, outline=stroke)
        elif pr[0] == "line":
            _, x1, y1, x2, y2, color = pr
            d.line([x1 * scale, y1 * scale, x2 * scale, y2 * scale], fill=color, width=scale)
            d.ellipse([(x2 - 2.5) * scale, (y2 - 2.5) * scale, (x2 + 2.5) * scale, (y2 + 2.5) * scale], fill=color)
        else:
            _, x, y, t, fill, anchor, size = pr
            f = font(size)
            t = _html.unescape(t)
            wpx = d.textlength(t, font=f)
            x0 = x * scale - (wpx / 2 if anchor == "middle" else 0)
            d.text((x0, (y - size) * scale), t, fill=fill, font=f)
    img.save(path)


def rank_row(s, y, cells, label=None, x0=70, cw=44, gap=130, labels_below=False):
    """One row of `len(cells)` ranks; cells[r] is a list of (color, text) slots held by rank r."""
    if label:
        s.text(8, y + 20, label, size=12)
    for r, slots in enumerate(cells):
        x = x0 + r * gap
        s.text(x + (cw * 0.62 * max(1, len(slots))) / 2, y + 44 if labels_below else y - 6, f"rank {r}", "#555", "middle", 11)
        if not slots:
            s.box(x, y, cw, 30, "white", "", GREY)
        for k, (c, t) in enumerate(slots):
            s.box(x + k * cw * 0.62, y, cw * 0.6, 30, c, t)


def before_after(name, title, before, after, arrows):
    n = len(before)
    s = Svg(90 + 130 * n, 215, title)
    s.text(8, 18, title, bold=True, size=14)
    rank_row(s, 50, before, "before")
    rank_row(s, 150, after, "after", labels_below=True)
    for a, b in arrows:
        s.arrow(70 + a * 130 + 20, 82, 70 + b * 130 + 20, 140)
    return s.save(name)


def slot(r, t=None):
    return (COL[r % 4], t if t is not None else f"t{r}")


def make_all():
    n = 4
    paths = []
    # --- point to point
    s = Svg(420, 130, "send / recv")
    s.text(8, 18, "send(tensor, dst=1) / recv(tensor, src=0)", bold=True, size=14)
    s.box(60, 50, 70, 40, COL[0], "rank 0")
    s.box(290, 50, 70, 40, COL[1], "rank 1")
    s.arrow(132, 70, 286, 70)
    s.text(210, 62, "tensor", "#333", "middle")
    s.text(210, 108, "CUDA tensors: ncclSend/ncclRecv over NVSwitch", "#666", "middle", 11)
    paths.append(s.save("send_recv"))
    # --- the six collectives
    paths.append(before_after("broadcast", "broadcast(tensor, src=0)", [[slot(0)], [], [], []],
                              [[slot(0)]] * n, [(0, r) for r in range(n)]))
    paths.append(before_after("scatter", "scatter(tensor, src=0, scatter_list)",
                              [[slot(k) for k in range(n)], [], [], []], [[slot(r)] for r in range(n)],
                              [(0, r) for r in range(n)]))
    paths.append(before_after("gather", "gather(tensor, dst=0, gather_list)", [[slot(r)] for r in range(n)],
                              [[slot(k) for k in range(n)], [], [], []], [(r, 0) for r in range(n)]))
    paths.append(before_after("reduce", "reduce(tensor, dst=0, op=SUM)", [[slot(r)] for r in range(n)],
                              [[("#333", "Σ")], [], [], []], [(r, 0) for r in range(n)]))
    paths.append(before_after("all_reduce", "all_reduce(tensor, op=SUM)", [[slot(r)] for r in range(n)],
                              [[("#333", "Σ")]] * n, [(a, b) for a in range(n) for b in range(n)]))
    paths.append(before_after("all_gather", "all_gather(tensor_list, tensor)", [[slot(r)] for r in range(n)],
                              [[slot(k) for k in range(n)]] * n, [(a, b) for a in range(n) for b in range(n)]))
    # --- ring all-reduce
    s = Svg(560, 240, "ring allreduce")
    s.text(8, 18, "allreduce(send, recv): ring on isend/recv, N-1 steps", bold=True, size=14)
    pos = [(120, 60), (380, 60), (380, 170), (120, 170)]
    for r, (x, y) in enumerate(pos):
        s.box(x, y, 80, 40, COL[r], f"rank {r}")
    for r in range(4):
        (x1, y1), (x2, y2) = pos[r], pos[(r + 1) % 4]
        if y1 == y2:
            s.arrow(x1 + (84 if x2 > x1 else -4), y1 + 20, x2 + (-4 if x2 > x1 else 84), y2 + 20)
        else:
            s.arrow(x1 + 40, y1 + (44 if y2 > y1 else -4), x2 + 40, y2 + (-4 if y2 > y1 else 44))
    s.text(280, 128, "accum += recv_buff;  send_buff &lt;-&gt; recv_buff", "#444", "middle", 12)
    s.text(280, 228, "chunked variant: reduce-scatter + all-gather, 2(N-1)/N x M bytes per rank", "#666", "middle", 11)
    paths.append(s.save("ring_allreduce"))
    # --- peer-memory variants
    s = Svg(760, 330, "peer memory all-reduce")
    s.text(8, 18, "fused all-reduce over symmetric peer memory (csrc/allreduce.cu)", bold=True, size=14)
    for r in range(4):
        s.box(40 + r * 180, 40, 140, 36, COL[r], f"GPU {r}: bucket[0..M)")
        s.arrow(110 + r * 180, 78, 300 + r * 54, 108, dash=True, color="#777")
    s.box(250, 110, 260, 34, "#333", "NVSwitch (NVLS: in-switch fp32 add)")
    rows = [("one-shot", "every GPU reads all N buckets (ld.global over NVLink), sums in rank order, scales by 1/N,",
             "writes its own copy: N x M bytes in, one barrier pair -- latency-optimal, small messages"),
            ("two-shot", "GPU r reduces slice r of every bucket, then stores the result into every peer's slice r:",
             "2 (N-1)/N x M bytes per GPU -- bandwidth-optimal without multicast"),
            ("NVLS", "multimem.ld_reduce on a multicast address: the switch adds the N copies in flight;",
             "multimem.st broadcasts the result -- M/N bytes read + M/N written per GPU")]
    for i, (name, l1, l2) in enumerate(rows):
        y = 176 + i * 44
        s.text(8, y, name, bold=True)
        s.text(90, y, l1, "#333", size=12)
        s.text(90, y + 16, l2, "#333", size=12)
    s.text(380, 318, "barrier = per-block flags in each buffer's signal pad, st.release.sys / ld.acquire.sys, "
           "monotonic epochs (CUDA-graph safe)", "#666", "middle", 11)
    paths.append(s.save("peer_allreduce"))
    # --- fused step
    s = Svg(900, 210, "fused training step")
    s.text(8, 18, "one ConvNet training step = one CUDA graph (ops/convnet_fused.py)", bold=True, size=14)
    boxes = [(20, 130, GREY, "H2D batch", "#222", ["uint8 pixels + labels from a", "pinned loader slot (copy stream)"]),
             (190, 250, COL[0], "convnet_step", "white", ["forward + loss + backward, one CTA (or a 2/4/8-CTA", "cluster) per sample; red.add.v4 into the flat bucket"]),
             (480, 250, COL[3], "allreduce_sgd", "white", ["push the bucket into the peers' inboxes (flag-in-data),", "sum in rank order, x 1/N, momentum SGD, re-zero"]),
             (770, 110, GREY, "D2H loss", "#222", ["running loss to a", "pinned word"])]
    for x, w, c, t, fg, desc in boxes:
        s.box(x, 50, w, 40, c, t, fg=fg)
        for k, line in enumerate(desc):
            s.text(x, 112 + 15 * k, line, "#555", size=11)
    for x in (152, 442, 732):
        s.arrow(x, 70, x + 36, 70)
    s.text(190, 165, "kernel 2 is pre-launched under kernel 1 with programmatic dependent launch (griddepcontrol); the exchange", "#666", size=11)
    s.text(190, 180, "crosses NVLink once: no barrier, no remote loads; gradient buckets and inboxes are double-buffered by step parity", "#666", size=11)
    paths.append(s.save("fused_step"))
    return paths


if __name__ == "__main__":
    for p in make_all():
        print("wrote", os.path.relpath(p, os.path.dirname(HERE)))
