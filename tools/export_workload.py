"""Write BASELINE.json config #2 (the dictionary and the query batch bench.py measures) as text files, one entry per line,
for the Go harness integration/go/b200_bench_test.go (the reference itself cannot run in this image).
usage: python tools/export_workload.py DIR [n_docs] [n_queries]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suggest_b200.workload import synthetic_workload, unpack  # noqa: E402


def main():
    out = sys.argv[1]
    n_docs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    n_queries = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    os.makedirs(out, exist_ok=True)
    docs, queries, _ = synthetic_workload(n_docs, n_queries)
    for name, packed in (("dictionary.txt", docs), ("queries.txt", queries)):
        with open(os.path.join(out, name), "wb") as f:
            f.write(b"\n".join(unpack(*packed)) + b"\n")
    print(f"{out}: {n_docs} dictionary entries, {n_queries} queries")


if __name__ == "__main__":
    main()
