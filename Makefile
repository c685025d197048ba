# Counterpart of the reference Makefile (targets `all` = docs, `ptp` = demo) plus build/test/bench.
.PHONY: all build test ptp demos bench docs clean

all: build docs

build:
	python -c "import __graft_entry__ as g; g.build()"

test:
	python -m pytest tests -x -q -m "not gpu"

ptp:
	python examples/gather_demo.py

demos: ptp
	python examples/p2p_demo.py
	python examples/groups_demo.py
	python examples/allreduce_demo.py --size 3

bench:
	python bench.py --gpus 1

docs:
	python docs/build_docs.py
	python docs/build_api.py
	python docs/build_pdf.py

clean:
	rm -rf dist_tuto.pth_b200/csrc/build dist_tuto.pth_b200/_C.so
