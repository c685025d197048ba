"""Docs toolchain (counterpart of the reference Makefile `all` target, which called an external paperify.py):
renders docs/tutorial.md to a standalone docs/tutorial.html with no external dependency, and (re)draws docs/figs/*.svg."""
import sys
import html
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def render(md: str) -> str:
    out, in_code, in_table, in_list = [], False, False, False
    for line in md.splitlines():
        if line.startswith("```"):
            out.append("</pre>" if in_code else "<pre>")
            in_code = not in_code
            continue
        if in_code:
            out.append(html.escape(line))
            continue
        if line.startswith("|"):
            cells = [c.strip() for c in line.strip().strip("|").split("|")]
            if all(set(c) <= set("-: ") for c in cells):
                continue
            if not in_table:
                out.append("<table border=1 cellpadding=4>")
                in_table = True
            out.append("<tr>" + "".join(f"<td>{inline(c)}</td>" for c in cells) + "</tr>")
            continue
        if in_table:
            out.append("</table>")
            in_table = False
        m = re.match(r"(#+) (.*)", line)
        if m:
            n = len(m.group(1))
            out.append(f"<h{n}>{inline(m.group(2))}</h{n}>")
        elif line.startswith("* "):
            if not in_list:
                out.append("<ul>")
                in_list = True
            out.append(f"<li>{inline(line[2:])}</li>")
        else:
            if in_list and not line.startswith("  "):
                out.append("</ul>")
                in_list = False
            out.append(f"<p>{inline(line)}</p>" if line.strip() else "")
    return "<html><head><meta charset='utf-8'><title>dist_tuto.pth_b200</title></head><body>" + "\n".join(out) + "</body></html>"


def inline(t: str) -> str:
    t = html.escape(t)
    t = re.sub(r"!\[([^\]]*)\]\(([^)]+)\)", r'<img alt="\1" src="\2" style="max-width:100%">', t)
    t = re.sub(r"`([^`]+)`", r"<code>\1</code>", t)
    t = re.sub(r"\*\*([^*]+)\*\*", r"<b>\1</b>", t)
    return re.sub(r"\*([^*]+)\*", r"<i>\1</i>", t)


if __name__ == "__main__":
    import make_figs   # docs/figs/*.svg (the reference ships its diagrams as binaries under figs/; ours are generated)

    make_figs.make_all()
    src = open(os.path.join(HERE, "tutorial.md")).read()
    page = render(src)
    for name in ("tutorial.html", "index.html"):       # the reference ships both tuto.html and index.html
        open(os.path.join(HERE, name), "w").write(page)
    print("wrote docs/tutorial.html, docs/index.html")
