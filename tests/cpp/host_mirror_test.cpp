// The reference's own end-to-end expectations, written the way its Go tests read, against the C++ host mirror.
//   pkg/suggest/ngram_index_test.go:15-40, example_test.go:14-72, service_test.go:11-80 (RAM driver, concurrent re-adds)
// usage: host_mirror_test <cars.dict>      (needs a B200; run by tests/test_cpp_mirror.py)
#include <cstdio>
#include <thread>

#include "suggest_b200.hpp"

#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

static const suggest::Dictionary kCollection = {"Nissan March", "Nissan Juke", "Nissan Maxima", "Nissan Murano",
                                                "Nissan Note", "Toyota Mark II", "Toyota Corolla", "Toyota Corona"};

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    {  // TestSuggestAuto
        suggest::IndexDescription d;
        d.Name = "index";
        auto index = suggest::NewRAMBuilder(kCollection, d)->Build();
        auto got = index->Suggest("Nissan ma", 0.5, metric::JaccardMetric(), 2);
        CHECK(got.size() == 2 && got[0].Key == 2 && got[1].Key == 0);
        CHECK(got[0].Score >= got[1].Score);
        // TestAutoComplete, ngram_index_test.go:42-67
        auto ac = index->Autocomplete("Niss", 5);
        CHECK(ac.size() == 5);
        for (uint32_t i = 0; i < 5; i++) CHECK(ac[i].Key == i);
        // every candidate of the T-occurrence count: the two returned above are among them with the overlaps their scores imply
        std::vector<uint32_t> size_a;
        auto all = index->Candidates({"Nissan ma"}, 0.5, metric::JaccardMetric(), &size_a);
        CHECK(size_a.size() == 1 && size_a[0] == 9);
        int seen = 0;
        for (const auto &c : all) {
            CHECK(c.Query == 0 && c.Overlap >= 1 && c.Overlap <= 9);
            const double score = (double)c.Overlap / (double)(9 + c.Segment - c.Overlap);
            CHECK(score >= 0.5);
            if (c.Position == 2 || c.Position == 0) seen++;
        }
        CHECK(seen == 2);
    }
    {  // Example
        suggest::IndexDescription d;
        d.Name = "cars";
        d.Alphabet = {"english", "$"};
        auto index = suggest::NewRAMBuilder(kCollection, d)->Build();
        auto got = index->Suggest("niss ma", 0.4, metric::CosineMetric(), 5);
        CHECK(got.size() == 2 && kCollection[got[0].Key] == "Nissan Maxima" && kCollection[got[1].Key] == "Nissan March");
    }
    {  // TestConcurrencyInMemory
        suggest::IndexDescription d;
        d.Name = "cars";
        d.Alphabet = {"russian", "english", "numbers", "$"};
        d.SourcePath = argv[1];
        auto service = suggest::NewService();
        service->AddRunTimeIndex(d);
        const char *words[5] = {"Nissan March", "Honda Fitt", "Wolfsvagen", "Tayota Corolla", "Micra Nissan"};
        const char *expected[5] = {"NISSAN MARCH", "HONDA FIT", nullptr, "TOYOTA COROLLA", "NISSAN MICRA"};
        int failures = 0;
        std::vector<std::thread> threads;
        for (int t = 0; t < 5; t++)
            threads.emplace_back([&] {
                for (int i = 0; i < 5; i++) {
                    auto res = service->Suggest("cars", suggest::NewSearchConfig(words[i], 5, metric::CosineMetric(), 0.7));
                    bool ok = expected[i] ? (res.size() == 1 && res[0].Value == expected[i]) : res.empty();
                    if (!ok) __atomic_fetch_add(&failures, 1, __ATOMIC_RELAXED);
                }
            });
        for (int t = 0; t < 3; t++) threads.emplace_back([&] { service->AddRunTimeIndex(d); });
        for (auto &t : threads) t.join();
        CHECK(failures == 0);
        bool threw = false;
        try {
            service->Suggest("nope", suggest::NewSearchConfig("x", 5, metric::CosineMetric(), 0.7));
        } catch (const suggest::Error &) { threw = true; }
        CHECK(threw);
        threw = false;
        try {
            suggest::NewSearchConfig("x", 0, metric::CosineMetric(), 0.7);
        } catch (const suggest::Error &) { threw = true; }
        CHECK(threw);
    }
    std::printf("host mirror ok, %llu kernel launches\n", (unsigned long long)sg_kernel_launches());
    return 0;
}
