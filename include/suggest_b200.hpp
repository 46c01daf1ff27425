// suggest_b200.hpp — header-only C++ host mirror of the reference's Go interfaces for the Suggest path, over the
// C ABI of libsuggest_b200.so (suggest_b200.h).  Same names, argument meaning and error behaviour as
//   pkg/suggest/config.go:25-35       IndexDescription
//   pkg/suggest/ngram_index_builder.go:14-57   Builder, NewRAMBuilder, NewFSBuilder
//   pkg/suggest/ngram_index.go:7-10, suggester.go:17-20   NGramIndex / Suggester
//   pkg/suggest/search.go:9-33        SearchConfig, NewSearchConfig
//   pkg/suggest/service.go:11-139     ResultItem, Service
//   pkg/metric/*.go                   JaccardMetric() ... ExactMetric()
// The reference is Go; this image has no Go toolchain, so this mirror (and the Python one) is what drives the library
// in tests.  Errors that Go returns as `error` are thrown as suggest::Error.
#pragma once
#include <cstdint>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "suggest_b200.h"

namespace metric {

// metric.Metric (pkg/metric/metric.go:7-16): the five built-ins are evaluated on the device; the object names one.
class Metric {
  public:
    explicit Metric(sg_metric code) : code_(code) {}
    sg_metric code() const { return code_; }

  private:
    sg_metric code_;
};
inline Metric JaccardMetric() { return Metric(SG_JACCARD); }
inline Metric CosineMetric() { return Metric(SG_COSINE); }
inline Metric DiceMetric() { return Metric(SG_DICE); }
inline Metric OverlapMetric() { return Metric(SG_OVERLAP); }
inline Metric ExactMetric() { return Metric(SG_EXACT); }

}  // namespace metric

namespace suggest {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

enum class Driver { RAM, DISC };  // config.go:16-23

struct IndexDescription {  // config.go:25-35
    Driver driver = Driver::RAM;
    std::string Name;
    int NGramSize = 3;
    std::string SourcePath, OutputPath;
    std::vector<std::string> Alphabet{"english", "russian", "numbers", "$"};
    std::string Pad = "$";
    std::string Wrap[2] = {"$", "$"};
    std::string basePath;
    int Device = 0;  // CUDA ordinal (not in the reference)

    std::string GetIndexPath() const { return !OutputPath.empty() && OutputPath[0] == '/' ? OutputPath : basePath + "/" + OutputPath; }
    std::string GetSourcePath() const { return !SourcePath.empty() && SourcePath[0] == '/' ? SourcePath : basePath + "/" + SourcePath; }
    std::string getHeaderFile() const { return Name + ".hd"; }
    std::string getDocumentListFile() const { return Name + ".dl"; }
};

struct Candidate {  // collector.go:12-17
    uint32_t Key;
    double Score;
};

struct ResultItem {  // service.go:11-16
    double Score;
    std::string Value;
};

class SearchConfig {  // search.go:9-15
  public:
    std::string query;
    int topK;
    metric::Metric metric;
    double similarity;
};

inline SearchConfig NewSearchConfig(const std::string &query, int topK, metric::Metric m, double similarity) {  // search.go:18-33
    if (topK <= 0) throw Error(SG_ERR_INVALID, "topK should be greater or equal to 1");
    if (similarity <= 0 || similarity > 1) throw Error(SG_ERR_INVALID, "similarity shouble be in (0.0, 1.0]");
    return SearchConfig{query, topK, m, similarity};
}

using Dictionary = std::vector<std::string>;  // dictionary.Dictionary: value by id

inline Dictionary OpenRAMDictionary(const std::string &path) {  // pkg/dictionary/helpers.go:25-48
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(SG_ERR_IO, "failed to open dictionary " + path);
    Dictionary d;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();  // bufio.ScanLines
        d.push_back(line);
    }
    return d;
}

// NGramIndex (Suggester part) backed by an sg_index handle
class NGramIndex {
  public:
    explicit NGramIndex(sg_index *h) : h_(h) {}
    ~NGramIndex() { sg_index_free(h_); }
    NGramIndex(const NGramIndex &) = delete;
    NGramIndex &operator=(const NGramIndex &) = delete;

    // nGramSuggester.Suggest with a FuzzyCollectorManager(topK), suggester.go:46-131
    std::vector<Candidate> Suggest(const std::string &query, double similarity, metric::Metric m, int topK) const {
        return SuggestBatch({query}, similarity, m, topK)[0];
    }

    std::vector<std::vector<Candidate>> SuggestBatch(const std::vector<std::string> &queries, double similarity, metric::Metric m,
                                                     int topK) const {
        std::string bytes;
        std::vector<uint32_t> off(1, 0);
        for (const auto &q : queries) {
            bytes += q;
            off.push_back((uint32_t)bytes.size());
        }
        const size_t n = queries.size(), k = topK > 0 ? (size_t)topK : 0;
        std::vector<uint32_t> ids(n * k + 1), counts(n + 1);
        std::vector<double> scores(n * k + 1);
        int rc = sg_search_batch(h_, bytes.data(), off.data(), (uint32_t)n, m.code(), similarity, (uint32_t)k, ids.data(),
                                 scores.data(), counts.data());
        if (rc != SG_OK) throw Error(rc, sg_last_error());
        std::vector<std::vector<Candidate>> out(n);
        for (size_t q = 0; q < n; q++)
            for (uint32_t i = 0; i < counts[q]; i++) out[q].push_back(Candidate{ids[q * k + i], scores[q * k + i]});
        return out;
    }

    // nGramAutocomplete.Autocomplete with a FirstKCollectorManager(limit), autocomplete.go:40-77
    std::vector<Candidate> Autocomplete(const std::string &query, int limit) const {
        const uint32_t off[2] = {0, (uint32_t)query.size()};
        const size_t k = limit > 0 ? (size_t)limit : 0;
        std::vector<uint32_t> ids(k + 1);
        std::vector<double> scores(k + 1);
        uint32_t count = 0;
        int rc = sg_autocomplete_batch(h_, query.data(), off, 1, (uint32_t)k, ids.data(), scores.data(), &count);
        if (rc != SG_OK) throw Error(rc, sg_last_error());
        std::vector<Candidate> out;
        for (uint32_t i = 0; i < count; i++) out.push_back(Candidate{ids[i], scores[i]});
        return out;
    }

    // What the mergers hand to Collector.Collect (pkg/merger/collector.go:10-13) over every admissible segment, for a
    // CollectorManager other than the fuzzy / first-k ones or a metric.Metric that is not built in: sg_candidates_batch.
    // `thresholds` = nullptr: the built-in metric m; otherwise [129][n_segments] bytes tabulated from the caller's metric.
    struct MergeCandidate {  // pkg/merger/list_merger.go:33-48, plus the query and the segment (sizeB) it came from
        uint32_t Query, Position, Overlap, Segment;
    };
    std::vector<MergeCandidate> Candidates(const std::vector<std::string> &queries, double similarity, metric::Metric m,
                                           std::vector<uint32_t> *sizeA = nullptr, const uint8_t *thresholds = nullptr) const {
        std::string bytes;
        std::vector<uint32_t> off(1, 0);
        for (const auto &q : queries) {
            bytes += q;
            off.push_back((uint32_t)bytes.size());
        }
        const size_t n = queries.size();
        std::vector<uint32_t> size_a(n + 1);
        uint64_t cap = 4 * n + 1024, total = 0;
        for (;;) {
            std::vector<uint32_t> q(cap), id(cap), ov(cap), seg(cap);
            int rc = sg_candidates_batch(h_, bytes.data(), off.data(), (uint32_t)n, m.code(), similarity, thresholds, cap, q.data(),
                                         id.data(), ov.data(), seg.data(), &total, size_a.data());
            if (rc != SG_OK) throw Error(rc, sg_last_error());
            if (total > cap) { cap = total; continue; }  // found more than the buffers hold: once more with the reported size
            std::vector<MergeCandidate> out(total);
            for (uint64_t i = 0; i < total; i++) out[i] = MergeCandidate{q[i], id[i], ov[i], seg[i]};
            if (sizeA) sizeA->assign(size_a.begin(), size_a.begin() + n);
            return out;
        }
    }

    sg_index_info Info() const {
        sg_index_info info{};
        sg_index_get_info(h_, &info);
        return info;
    }
    sg_index *handle() const { return h_; }

  private:
    sg_index *h_;
};

class Builder {  // ngram_index_builder.go:14-17
  public:
    virtual ~Builder() = default;
    virtual std::shared_ptr<NGramIndex> Build() = 0;
};

namespace detail {
struct CConfig {
    sg_config cfg{};
    std::vector<const char *> alpha;
    explicit CConfig(const IndexDescription &d) {
        for (const auto &a : d.Alphabet) alpha.push_back(a.c_str());
        cfg.ngram_size = d.NGramSize;
        cfg.wrap_start = d.Wrap[0].c_str();
        cfg.wrap_end = d.Wrap[1].c_str();
        cfg.pad = d.Pad.c_str();
        cfg.alphabet = alpha.data();
        cfg.n_alphabet = (int32_t)alpha.size();
        cfg.device = d.Device;
    }
};
}  // namespace detail

class RAMBuilder : public Builder {  // ngram_index_builder.go:27-35
  public:
    RAMBuilder(const Dictionary &dict, IndexDescription d, uint32_t id_base = 0) : dict_(dict), d_(std::move(d)), id_base_(id_base) {}
    std::shared_ptr<NGramIndex> Build() override {
        std::string bytes;
        std::vector<uint64_t> off(1, 0);
        for (const auto &v : dict_) {
            bytes += v;
            off.push_back(bytes.size());
        }
        detail::CConfig c(d_);
        sg_index *h = nullptr;
        int rc = sg_index_build(&c.cfg, bytes.data(), off.data(), (uint32_t)dict_.size(), id_base_, &h);
        if (rc != SG_OK) throw Error(rc, std::string("failed to build NGramIndex: ") + sg_last_error());
        return std::make_shared<NGramIndex>(h);
    }

  private:
    const Dictionary &dict_;
    IndexDescription d_;
    uint32_t id_base_;
};

class FSBuilder : public Builder {  // ngram_index_builder.go:38-57
  public:
    explicit FSBuilder(IndexDescription d) : d_(std::move(d)) {}
    std::shared_ptr<NGramIndex> Build() override {
        detail::CConfig c(d_);
        sg_index *h = nullptr;
        const std::string hd = d_.GetIndexPath() + "/" + d_.getHeaderFile(), dl = d_.GetIndexPath() + "/" + d_.getDocumentListFile();
        int rc = sg_index_open_disk(&c.cfg, hd.c_str(), dl.c_str(), &h);
        if (rc != SG_OK) throw Error(rc, std::string("failed to open FS inverted index: ") + sg_last_error());
        return std::make_shared<NGramIndex>(h);
    }

  private:
    IndexDescription d_;
};

inline std::unique_ptr<Builder> NewRAMBuilder(const Dictionary &dict, const IndexDescription &d) {
    return std::unique_ptr<Builder>(new RAMBuilder(dict, d));
}
inline std::unique_ptr<Builder> NewFSBuilder(const IndexDescription &d) { return std::unique_ptr<Builder>(new FSBuilder(d)); }

class Service {  // service.go:18-139
  public:
    void AddIndexByDescription(const IndexDescription &d) { d.driver == Driver::RAM ? AddRunTimeIndex(d) : AddOnDiscIndex(d); }

    void AddRunTimeIndex(const IndexDescription &d) {
        auto dict = std::make_shared<Dictionary>(OpenRAMDictionary(d.GetSourcePath()));
        RAMBuilder b(*dict, d);
        AddIndex(d.Name, dict, b);
    }

    // the values come from the description's source file: the CDB reader stays on the Go side
    void AddOnDiscIndex(const IndexDescription &d) {
        auto dict = std::make_shared<Dictionary>(OpenRAMDictionary(d.GetSourcePath()));
        FSBuilder b(d);
        AddIndex(d.Name, dict, b);
    }

    void AddIndex(const std::string &name, std::shared_ptr<Dictionary> dict, Builder &builder) {
        std::shared_ptr<NGramIndex> index;
        try {
            index = builder.Build();
        } catch (const Error &e) {
            throw Error(e.code, std::string("failed to build NGramIndex: ") + e.what());
        }
        std::unique_lock<std::shared_mutex> lk(mu_);
        indexes_[name] = std::move(index);  // a replaced index lives until its searches finish (shared_ptr)
        dictionaries_[name] = std::move(dict);
    }

    std::vector<std::string> GetDictionaries() const {
        std::shared_lock<std::shared_mutex> lk(mu_);
        std::vector<std::string> names;
        for (const auto &kv : dictionaries_) names.push_back(kv.first);
        return names;
    }

    std::vector<ResultItem> Suggest(const std::string &dictName, const SearchConfig &config) const {
        std::shared_ptr<NGramIndex> index;
        std::shared_ptr<Dictionary> dict;
        {
            std::shared_lock<std::shared_mutex> lk(mu_);
            auto i = indexes_.find(dictName);
            auto d = dictionaries_.find(dictName);
            if (i == indexes_.end() || d == dictionaries_.end())
                throw Error(SG_ERR_INVALID, "given dictionary " + dictName + " is not exists");
            index = i->second;
            dict = d->second;
        }
        std::vector<ResultItem> result;
        for (const Candidate &c : index->Suggest(config.query, config.similarity, config.metric, config.topK))
            result.push_back(ResultItem{c.Score, dict->at(c.Key)});
        return result;
    }

    // Service.Autocomplete, service.go:141-172 (score 0 for every item, as the reference reports it)
    std::vector<ResultItem> Autocomplete(const std::string &dictName, const std::string &query, int limit) const {
        std::shared_ptr<NGramIndex> index;
        std::shared_ptr<Dictionary> dict;
        {
            std::shared_lock<std::shared_mutex> lk(mu_);
            auto i = indexes_.find(dictName);
            auto d = dictionaries_.find(dictName);
            if (i == indexes_.end() || d == dictionaries_.end())
                throw Error(SG_ERR_INVALID, "given dictionary " + dictName + " is not exists");
            index = i->second;
            dict = d->second;
        }
        std::vector<ResultItem> result;
        for (const Candidate &c : index->Autocomplete(query, limit)) result.push_back(ResultItem{0.0, dict->at(c.Key)});
        return result;
    }

  private:
    mutable std::shared_mutex mu_;
    std::map<std::string, std::shared_ptr<NGramIndex>> indexes_;
    std::map<std::string, std::shared_ptr<Dictionary>> dictionaries_;
};

inline std::unique_ptr<Service> NewService() { return std::unique_ptr<Service>(new Service()); }

}  // namespace suggest
