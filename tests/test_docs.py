"""Docs toolchain (reference Makefile `all` target): figures are well-formed SVG, the tutorial renders with them."""
import os
import sys
import xml.dom.minidom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "docs"))


def test_figures_are_wellformed_svg(tmp_path, monkeypatch):
    import make_figs
    monkeypatch.setattr(make_figs, "OUT", str(tmp_path))
    paths = make_figs.make_all()
    names = {os.path.basename(p) for p in paths}
    # the reference's figs/ set (send_recv, broadcast, scatter, gather, reduce, all_reduce, all_gather) + ours
    for want in ("send_recv", "broadcast", "scatter", "gather", "reduce", "all_reduce", "all_gather", "ring_allreduce",
                 "peer_allreduce", "fused_step"):
        assert want + ".svg" in names
    for p in paths:
        doc = xml.dom.minidom.parse(p)
        assert doc.documentElement.tagName == "svg"


def test_tutorial_renders_with_figures():
    import build_docs
    md = open(os.path.join(ROOT, "docs", "tutorial.md")).read()
    html = build_docs.render(md)
    assert html.count("<img") >= 10 and "<h2>" in html and "<pre>" in html
    for sec in ("Setup", "Point-to-Point Communication", "Collective Communication", "Distributed Training",
                "Our Own Ring-Allreduce", "Communication Backends", "Initialization Methods"):
        assert sec in html          # the section structure of tuto.md


def test_api_reference_lists_every_tutorial_name(tmp_path, monkeypatch):
    import build_api
    monkeypatch.setattr(build_api, "HERE", str(tmp_path))
    build_api.main()
    text = open(tmp_path / "api.md").read()
    # the API surface the tutorial documents
    for name in ("init_processes", "send", "recv", "isend", "irecv", "new_group", "all_reduce", "reduce", "broadcast", "scatter",
                 "gather", "all_gather", "reduce_op", "allreduce", "Partition", "DataPartitioner", "partition_dataset", "Net",
                 "average_gradients", "run"):
        assert f"`{name}" in text, name


def test_pdf_is_structurally_valid():
    import re
    import build_pdf
    pdf = build_pdf.build(open(os.path.join(ROOT, "docs", "tutorial.md")).read())
    assert pdf.startswith(b"%PDF-1.4") and pdf.rstrip().endswith(b"%%EOF")
    sx = int(re.search(rb"startxref\n(\d+)", pdf).group(1))
    assert pdf[sx:sx + 4] == b"xref"
    offsets = [int(o) for o in re.findall(rb"(\d{10}) 00000 n", pdf[sx:])]
    for i, o in enumerate(offsets, 1):                      # every xref entry points at its object
        assert pdf[o:].startswith(b"%d 0 obj" % i)
    for m in re.finditer(rb"<< /Length (\d+) >>\nstream\n", pdf):   # stream lengths are exact
        end = m.end() + int(m.group(1))
        assert pdf[end:end + 10] == b"\nendstream"
    assert pdf.count(b"/Type /Page /Parent") >= 3 and b"Distributed Training" in pdf
