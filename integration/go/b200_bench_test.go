//go:build cgo
// +build cgo

// b200_bench_test.go - times the reference's own Suggest path (and, built with cgo, the B200 index) on the workload
// bench.py measures: BASELINE.json config #2, exported by `python tools/export_workload.py DIR`.
// Goes next to b200.go in github.com/suggest-go/suggest/pkg/suggest.  NOT RUN in this repository's build image (no Go
// toolchain); bench.py's reference arm times the line-faithful C restatement (oracle/) instead and says so.
//
//   SUGGEST_WORKLOAD=DIR go test ./pkg/suggest -run xxx -bench 'Config2' -cpu 16
package suggest

import (
	"bufio"
	"os"
	"sync/atomic"
	"testing"
	"time"

	"github.com/suggest-go/suggest/pkg/dictionary"
	"github.com/suggest-go/suggest/pkg/metric"
)

func readLines(b *testing.B, path string) []string {
	f, err := os.Open(path)
	if err != nil {
		b.Skipf("workload not exported (%v): python tools/export_workload.py DIR, then SUGGEST_WORKLOAD=DIR", err)
	}
	defer f.Close()
	var lines []string
	sc := bufio.NewScanner(f)
	for sc.Scan() {
		lines = append(lines, sc.Text())
	}
	return lines
}

func config2(b *testing.B) (dictionary.Dictionary, []string, IndexDescription) {
	dir := os.Getenv("SUGGEST_WORKLOAD")
	if dir == "" {
		b.Skip("SUGGEST_WORKLOAD is not set")
	}
	description := IndexDescription{ // pkg/suggest/ngram_index_test.go:216-223
		Driver: RAMDriver, Name: "index", NGramSize: 3, Pad: "$", Wrap: [2]string{"$", "$"},
		Alphabet: []string{"english", "russian", "numbers", "$"},
	}
	return dictionary.NewInMemoryDictionary(readLines(b, dir+"/dictionary.txt")), readLines(b, dir+"/queries.txt"), description
}

func runConfig2(b *testing.B, index NGramIndex, queries []string) {
	var next uint64
	b.ReportAllocs()
	b.ResetTimer()
	start := time.Now() // (testing.B.Elapsed needs go 1.20; go.mod says 1.13)
	b.RunParallel(func(pb *testing.PB) { // one query per goroutine at a time, as the HTTP server runs them
		for pb.Next() {
			q := queries[int(atomic.AddUint64(&next, 1))%len(queries)]
			if _, err := index.Suggest(q, 0.5, metric.JaccardMetric(), newFuzzyCollectorManager(10)); err != nil {
				b.Fatal(err)
			}
		}
	})
	b.ReportMetric(float64(b.N)/time.Since(start).Seconds(), "queries/s")
}

// BenchmarkConfig2Reference: the unmodified reference (RAM driver, CPMerge, five goroutines per query).
func BenchmarkConfig2Reference(b *testing.B) {
	dict, queries, description := config2(b)
	builder, err := NewRAMBuilder(dict, description)
	if err != nil {
		b.Fatal(err)
	}
	index, err := builder.Build()
	if err != nil {
		b.Fatal(err)
	}
	runConfig2(b, index, queries)
}

// BenchmarkConfig2B200Batch: the same queries through b200.go's SuggestBatch, 65,536 per call.
func BenchmarkConfig2B200Batch(b *testing.B) {
	dict, queries, description := config2(b)
	index, err := NewB200Builder(dict, description, 0).Build()
	if err != nil {
		b.Skipf("no B200 index: %v", err)
	}
	batched, ok := index.(*nGramIndex).suggester.(*b200Index)
	if !ok {
		b.Fatal("NewB200Builder did not return a b200Index")
	}
	b.ResetTimer()
	start := time.Now()
	done := 0
	for i := 0; i < b.N; i++ {
		res, err := batched.SuggestBatch(queries, 0.5, metric.JaccardMetric(), newFuzzyCollectorManager(10))
		if err != nil {
			b.Fatal(err)
		}
		done += len(res)
	}
	b.ReportMetric(float64(done)/time.Since(start).Seconds(), "queries/s")
}
