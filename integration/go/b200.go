//go:build cgo
// +build cgo

// b200.go - the one file a maintainer adds to github.com/suggest-go/suggest/pkg/suggest to serve NGramIndex from
// libsuggest_b200.so (include/suggest_b200.h).  It lives in package suggest because topKQueue.topK and
// FirstKCollectorManager.limit are unexported, exactly as suggester.go:101 already type-switches on *FuzzyCollectorManager.
//
// NOT COMPILED in the build image of this repository (no Go toolchain there).  The same marshalling is exercised through
// ctypes (suggest_b200/suggest.py) and C++ (include/suggest_b200.hpp); see INTEGRATION.md.
//
// Wiring:  service.AddIndex(name, dict, suggest.NewB200Builder(dict, description, 0))   instead of AddRunTimeIndex /
// AddOnDiscIndex (pkg/suggest/service.go:52-91).  Nothing else in the reference changes.
package suggest

/*
#cgo LDFLAGS: -lsuggest_b200
#include <stdlib.h>
#include "suggest_b200.h"
*/
import "C"

import (
	"errors"
	"reflect"
	"runtime"
	"sort"
	"sync"
	"unsafe"

	"github.com/suggest-go/suggest/pkg/dictionary"
	"github.com/suggest-go/suggest/pkg/index"
	"github.com/suggest-go/suggest/pkg/merger"
	"github.com/suggest-go/suggest/pkg/metric"
	"github.com/suggest-go/suggest/pkg/utils"
)

// NewB200Builder builds the index on the GPU from the dictionary (RAM driver) or opens <name>.hd / <name>.dl (DISC driver).
func NewB200Builder(dict dictionary.Dictionary, description IndexDescription, device int) Builder {
	return &b200Builder{dict: dict, description: description, device: device}
}

type b200Builder struct {
	dict        dictionary.Dictionary
	description IndexDescription
	device      int
}

func lastError() error { return errors.New(C.GoString(C.sg_last_error())) }

// config marshals the tokenizer-relevant fields of IndexDescription (config.go:25-35) into C memory.
func (b *b200Builder) config() (C.sg_config, func()) {
	d := b.description
	n := len(d.Alphabet)
	pa := (**C.char)(C.malloc(C.size_t(n+1) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	alphabet := (*[1 << 20]*C.char)(unsafe.Pointer(pa))[: n+1 : n+1] // go.mod says go 1.13: no unsafe.Slice
	for i, a := range d.Alphabet {
		alphabet[i] = C.CString(a)
	}
	cfg := C.sg_config{
		ngram_size: C.int32_t(d.NGramSize), wrap_start: C.CString(d.Wrap[0]), wrap_end: C.CString(d.Wrap[1]),
		pad: C.CString(d.Pad), alphabet: pa, n_alphabet: C.int32_t(n), device: C.int32_t(b.device),
	}
	return cfg, func() {
		for i := 0; i < n; i++ {
			C.free(unsafe.Pointer(alphabet[i]))
		}
		C.free(unsafe.Pointer(pa))
		C.free(unsafe.Pointer(cfg.wrap_start))
		C.free(unsafe.Pointer(cfg.wrap_end))
		C.free(unsafe.Pointer(cfg.pad))
	}
}

// Build implements Builder (ngram_index_builder.go:14-17).
func (b *b200Builder) Build() (NGramIndex, error) {
	cfg, release := b.config()
	defer release()
	var handle *C.sg_index
	if b.description.Driver == DiscDriver {
		hd := C.CString(b.description.GetIndexPath() + "/" + b.description.getHeaderFile())
		dl := C.CString(b.description.GetIndexPath() + "/" + b.description.getDocumentListFile())
		defer C.free(unsafe.Pointer(hd))
		defer C.free(unsafe.Pointer(dl))
		if rc := C.sg_index_open_disk(&cfg, hd, dl, &handle); rc != 0 {
			return nil, lastError()
		}
	} else {
		var bytes []byte
		offsets := []C.uint64_t{0}
		// ids are dense line numbers (pkg/dictionary/helpers.go:38-45), so the i-th value is document i
		err := b.dict.Iterate(func(key dictionary.Key, value dictionary.Value) error {
			bytes = append(bytes, value...)
			offsets = append(offsets, C.uint64_t(len(bytes)))
			return nil
		})
		if err != nil {
			return nil, err
		}
		if rc := C.sg_index_build(&cfg, bytesPtr(bytes), &offsets[0], C.uint32_t(len(offsets)-1), 0, &handle); rc != 0 {
			return nil, lastError()
		}
	}
	ix := &b200Index{handle: handle}
	// Service.Suggest is called with one query per goroutine (internal/suggest/api/suggest_handler.go:56): the batcher
	// coalesces those calls into sg_search_batch calls over page-locked buffers of its own
	if rc := C.sg_batcher_create(handle, batcherMaxBatch, batcherMaxWaitUs, batcherMaxK, &ix.batcher); rc != 0 {
		C.sg_index_free(handle)
		return nil, lastError()
	}
	runtime.SetFinalizer(ix, func(i *b200Index) { // as index_reader.go:49-51 does for mmaps
		C.sg_batcher_free(i.batcher)
		C.sg_index_free(i.handle)
	})
	return NewNGramIndex(ix, ix), nil
}

const (
	batcherMaxBatch  = 16384
	batcherMaxWaitUs = 100
	batcherMaxK      = 256
)

type b200Index struct {
	handle  *C.sg_index
	batcher *C.sg_batcher
}

// pinnedRows are page-locked result rows (sg_pinned_alloc) of sg_candidate entries - 16 bytes, {key uint32, pad, score
// float64}, which is exactly how Candidate{Key index.Position; Score float64} lies in memory (collector.go:12-17) - plus the
// counts.  sg_search_batch_candidates lets the kernels store straight into them (a Go-heap slice is pageable and would be
// staged), one PCIe write per candidate.  Pooled, so that a batch call costs no cudaHostAlloc.
type pinnedRows struct {
	n, k   int
	rows   unsafe.Pointer
	counts unsafe.Pointer
}

var pinnedPool sync.Pool

func getPinnedRows(n, k int) (*pinnedRows, error) {
	if v := pinnedPool.Get(); v != nil {
		r := v.(*pinnedRows)
		if r.n >= n && r.n*r.k >= n*k {
			return r, nil
		}
		r.free()
	}
	r := &pinnedRows{n: n, k: k}
	if C.sg_pinned_alloc(C.uint64_t(n*k)*C.uint64_t(unsafe.Sizeof(C.sg_candidate{})), &r.rows) != 0 ||
		C.sg_pinned_alloc(C.uint64_t(n*4), &r.counts) != 0 {
		r.free()
		return nil, lastError()
	}
	runtime.SetFinalizer(r, func(r *pinnedRows) { r.free() })
	return r, nil
}

func (r *pinnedRows) free() {
	C.sg_pinned_free(r.rows)
	C.sg_pinned_free(r.counts)
	r.rows, r.counts = nil, nil
}

func bytesPtr(b []byte) *C.char {
	if len(b) == 0 {
		return nil
	}
	return (*C.char)(unsafe.Pointer(&b[0]))
}

func pack(queries []string) ([]byte, []C.uint32_t) {
	var bytes []byte
	offsets := make([]C.uint32_t, 1, len(queries)+1)
	for _, q := range queries {
		bytes = append(bytes, q...)
		offsets = append(offsets, C.uint32_t(len(bytes)))
	}
	return bytes, offsets
}

// metricCode: the five built-ins are evaluated on the device; their concrete types are unexported, compare dynamic types.
func metricCode(m metric.Metric) (C.int, bool) {
	same := func(a, b metric.Metric) bool { return reflect.TypeOf(a) == reflect.TypeOf(b) }
	switch {
	case same(m, metric.JaccardMetric()):
		return C.SG_JACCARD, true
	case same(m, metric.CosineMetric()):
		return C.SG_COSINE, true
	case same(m, metric.DiceMetric()):
		return C.SG_DICE, true
	case same(m, metric.OverlapMetric()):
		return C.SG_OVERLAP, true
	case same(m, metric.ExactMetric()):
		return C.SG_EXACT, true
	}
	return 0, false
}

func rowsToCandidates(n, k int, ids []C.uint32_t, scores []C.double, counts []C.uint32_t) [][]Candidate {
	out := make([][]Candidate, n)
	for q := 0; q < n; q++ {
		row := make([]Candidate, counts[q])
		for i := range row {
			row[i] = Candidate{Key: index.Position(ids[q*k+i]), Score: float64(scores[q*k+i])}
		}
		out[q] = row
	}
	return out
}

// Suggest implements Suggester (suggester.go:46-131).  The fuzzy top-k with a built-in metric goes through the batcher:
// concurrent callers (one goroutine per HTTP request) share sg_search_batch calls; everything else is a batch of one.
func (ix *b200Index) Suggest(query string, similarity float64, m metric.Metric, factory CollectorManagerFactory) ([]Candidate, error) {
	if fuzzy, ok := factory().(*FuzzyCollectorManager); ok {
		if queue, ok := fuzzy.globalQueue.(*topKQueue); ok && queue.topK > 0 && queue.topK <= batcherMaxK {
			if code, known := metricCode(m); known {
				k := queue.topK
				ids := make([]C.uint32_t, k)
				scores := make([]C.double, k)
				var count C.uint32_t
				q := []byte(query)
				rc := C.sg_suggest_one(ix.batcher, bytesPtr(q), C.uint32_t(len(q)), code, C.double(similarity), C.uint32_t(k),
					&ids[0], &scores[0], &count)
				runtime.KeepAlive(ix)
				if rc != 0 {
					return nil, lastError()
				}
				return rowsToCandidates(1, k, ids, scores, []C.uint32_t{count})[0], nil
			}
		}
	}
	res, err := ix.SuggestBatch([]string{query}, similarity, m, factory)
	if err != nil {
		return nil, err
	}
	return res[0], nil
}

// SuggestBatch is the additive batched call (the reference has none); put a micro-batcher behind Suggest to use it from
// the HTTP handlers (one goroutine per request; a cgo call blocks an OS thread, so large batches are what Go wants too).
func (ix *b200Index) SuggestBatch(queries []string, similarity float64, m metric.Metric, factory CollectorManagerFactory) ([][]Candidate, error) {
	fuzzy, ok := factory().(*FuzzyCollectorManager)
	code, known := metricCode(m)
	if !ok || !known {
		return ix.suggestReplay(queries, similarity, m, factory)
	}
	queue, ok := fuzzy.globalQueue.(*topKQueue)
	if !ok {
		return ix.suggestReplay(queries, similarity, m, factory)
	}
	k := queue.topK
	bytes, offsets := pack(queries)
	n := len(queries)
	if n == 0 || k <= 0 {
		return make([][]Candidate, n), nil
	}
	// page-locked rows from the pool: the kernels store the valid entries of every row straight into them, laid out as
	// []Candidate; a row is copied out with one copy(), no per-entry conversion
	rows, err := getPinnedRows(n, k)
	if err != nil {
		return nil, err
	}
	defer pinnedPool.Put(rows)
	if unsafe.Sizeof(Candidate{}) != unsafe.Sizeof(C.sg_candidate{}) || unsafe.Offsetof(Candidate{}.Score) != 8 {
		return nil, errors.New("b200 index: Candidate is not laid out as sg_candidate")
	}
	rc := C.sg_search_batch_candidates(ix.handle, bytesPtr(bytes), &offsets[0], C.uint32_t(n), code, C.double(similarity), C.uint32_t(k),
		(*C.sg_candidate)(rows.rows), (*C.uint32_t)(rows.counts))
	runtime.KeepAlive(ix)
	if rc != 0 {
		return nil, lastError()
	}
	view := (*[1 << 26]Candidate)(rows.rows)[: n*k : n*k]
	counts := (*[1 << 28]C.uint32_t)(rows.counts)[:n:n]
	out := make([][]Candidate, n)
	for q := 0; q < n; q++ {
		out[q] = append([]Candidate(nil), view[q*k:q*k+int(counts[q])]...)
	}
	return out, nil
}

// Autocomplete implements Autocomplete (autocomplete.go:40-77).  FirstKCollectorManager(limit) runs on the device
// (sg_autocomplete_batch: T = len(tokens) over segments len(tokens)..S-1, no tail wrap, lowest ids win, score -id).
func (ix *b200Index) Autocomplete(query string, factory CollectorManagerFactory) ([]Candidate, error) {
	first, ok := factory().(*FirstKCollectorManager)
	if !ok {
		return nil, errors.New("b200 index: Autocomplete serves FirstKCollectorManager (the spellchecker's collector: sg_predict_batch)")
	}
	k := first.limit
	if k <= 0 {
		return []Candidate{}, nil
	}
	bytes, offsets := pack([]string{query})
	ids := make([]C.uint32_t, k)
	scores := make([]C.double, k)
	counts := make([]C.uint32_t, 1)
	rc := C.sg_autocomplete_batch(ix.handle, bytesPtr(bytes), &offsets[0], 1, C.uint32_t(k), &ids[0], &scores[0], &counts[0])
	runtime.KeepAlive(ix)
	if rc != 0 {
		return nil, lastError()
	}
	return rowsToCandidates(1, k, ids, scores, counts)[0], nil
}

// suggestReplay serves the two interface wrinkles: a CollectorManager that is not the fuzzy one, and a metric.Metric that
// is not built in.  The device returns every candidate of the T-occurrence count (sg_candidates_batch) and the caller's
// manager is driven exactly as suggester.go:66-108 drives it.
func (ix *b200Index) suggestReplay(queries []string, similarity float64, m metric.Metric, factory CollectorManagerFactory) ([][]Candidate, error) {
	var info C.sg_index_info
	if rc := C.sg_index_get_info(ix.handle, &info); rc != 0 {
		return nil, lastError()
	}
	S := int(info.n_segments)
	n := len(queries)
	if n == 0 {
		return [][]Candidate{}, nil
	}
	var table []C.uint8_t
	code, known := metricCode(m)
	if !known { // tabulate the caller's metric: table[a*S+B] = Threshold over its window, 0 elsewhere
		table = make([]C.uint8_t, (C.SG_MAX_QUERY_TOKENS+1)*S)
		for a := 1; a <= C.SG_MAX_QUERY_TOKENS; a++ {
			for B := utils.Max(m.MinY(similarity, a), 0); B <= utils.Min(m.MaxY(similarity, a), S-1); B++ {
				table[a*S+B] = C.uint8_t(utils.Min(utils.Max(m.Threshold(similarity, a, B), 0), 255))
			}
		}
	}
	bytes, offsets := pack(queries)
	sizeA := make([]C.uint32_t, n)
	capacity := 4*n + 1024
	var cq, cid, cov, cseg []C.uint32_t
	var total C.uint64_t
	for {
		cq, cid = make([]C.uint32_t, capacity), make([]C.uint32_t, capacity)
		cov, cseg = make([]C.uint32_t, capacity), make([]C.uint32_t, capacity)
		var thr *C.uint8_t
		if table != nil {
			thr = &table[0]
		}
		rc := C.sg_candidates_batch(ix.handle, bytesPtr(bytes), &offsets[0], C.uint32_t(n), code, C.double(similarity), thr,
			C.uint64_t(capacity), &cq[0], &cid[0], &cov[0], &cseg[0], &total, &sizeA[0])
		runtime.KeepAlive(ix)
		if rc != 0 {
			return nil, lastError()
		}
		if int(total) <= capacity {
			break
		}
		capacity = int(total) // found more than the buffers hold: once more with the reported size
	}
	// per query, per segment, positions ascending: the order the mergers emit
	order := make([]int, int(total))
	for i := range order {
		order[i] = i
	}
	sort.Slice(order, func(x, y int) bool {
		a, b := order[x], order[y]
		if cq[a] != cq[b] {
			return cq[a] < cq[b]
		}
		if cseg[a] != cseg[b] {
			return cseg[a] < cseg[b]
		}
		return cid[a] < cid[b]
	})
	perQuery := make([]map[int][]merger.MergeCandidate, n)
	for _, i := range order {
		q := int(cq[i])
		if perQuery[q] == nil {
			perQuery[q] = map[int][]merger.MergeCandidate{}
		}
		perQuery[q][int(cseg[i])] = append(perQuery[q][int(cseg[i])], merger.NewMergeCandidate(uint32(cid[i]), uint32(cov[i])))
	}
	out := make([][]Candidate, n)
	for i := range queries {
		a := int(sizeA[i])
		if a == 0 {
			out[i] = []Candidate{} // suggester.go:49-51
			continue
		}
		manager := factory()
		bMin, bMax := m.MinY(similarity, a), utils.Min(m.MaxY(similarity, a), S-1)
		feed := make([]int, 0, 2*(bMax-bMin+2))
		for x, y := a, a+1; x >= bMin || y <= bMax; x, y = x-1, y+1 { // feed order of suggester.go:110-118
			if x >= bMin {
				feed = append(feed, x)
			}
			if y <= bMax {
				feed = append(feed, y)
			}
		}
		for _, sizeB := range feed {
			if sizeB < 0 || sizeB >= S {
				continue // indices.Get(sizeB) == nil
			}
			if t := m.Threshold(similarity, a, sizeB); t <= 0 || t > sizeB || t > a {
				continue
			}
			collector := manager.Create()
			collector.SetScorer(NewMetricScorer(m, a, sizeB))
			for _, c := range perQuery[i][sizeB] {
				if err := collector.Collect(c); err != nil {
					if err == merger.ErrCollectionTerminated {
						break
					}
					return nil, err
				}
			}
			if err := manager.Collect(collector); err != nil {
				return nil, err
			}
		}
		out[i] = manager.GetCandidates()
	}
	return out, nil
}
