// fuzz_parsers.cpp -- mutation fuzzing of the byte-stream parsers that face files from outside:
// the H.264 Annex-B indexer, the mp4 demuxer / muxer and the stored-descriptor (proto3 wire) readers.
// Built with AddressSanitizer + UBSan by tests/test_storage_cpu.py; every mutated input lives in an
// exact-size heap block so that an over-read of a single byte is reported.
//   fuzz_parsers <seed> <iterations> <file>...      (file kind by content: mp4, Annex-B, else descriptor)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>

#include "h264.h"
#include "mp4.h"
#include "storage.h"

using namespace scanner;
using namespace scanner::internal;

static std::vector<u8> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<u8>((std::istreambuf_iterator<char>(f)), {});
}

static std::vector<u8> mutate(const std::vector<u8>& seed, std::mt19937_64& rng) {
  std::vector<u8> m = seed;
  switch (rng() % 4) {
    case 0:  // scattered byte flips
      for (int c = 1 + (int)(rng() % 8); c > 0; --c) m[rng() % m.size()] = (u8)rng();
      break;
    case 1:  // truncation
      m.resize(rng() % m.size());
      break;
    case 2: {  // one 32-bit field
      const size_t i = rng() % m.size();
      for (size_t j = 0; j < 4 && i + j < m.size(); ++j) m[i + j] = (u8)rng();
      break;
    }
    default:  // headers live at the two ends of a file
      for (int c = 1 + (int)(rng() % 6); c > 0; --c) {
        const size_t lim = std::min<size_t>(700, m.size());
        m[(rng() & 1) ? rng() % lim : m.size() - 1 - rng() % lim] = (u8)rng();
      }
  }
  std::vector<u8> exact(m.begin(), m.end());
  exact.shrink_to_fit();
  return exact;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::mt19937_64 rng((unsigned long long)atoll(argv[1]));
  const int iters = atoi(argv[2]);
  long accepted = 0, rejected = 0;
  for (int a = 3; a < argc; ++a) {
    const std::vector<u8> seed = slurp(argv[a]);
    if (seed.empty()) {
      fprintf(stderr, "empty seed %s\n", argv[a]);
      return 2;
    }
    const bool mp4 = looks_like_mp4(seed.data(), seed.size());
    const bool annexb = !mp4 && seed.size() > 4 && seed[0] == 0 && seed[1] == 0 && (seed[2] == 1 || (seed[2] == 0 && seed[3] == 1));
    for (int it = 0; it < iters; ++it) {
      const std::vector<u8> m = mutate(seed, rng);
      bool ok = false;
      if (mp4) {
        Mp4Track t;
        ok = demux_mp4(m.data(), m.size(), t).success();
        if (ok) {
          H264Index idx;
          index_bytestream(t.annexb.data(), t.annexb.size(), idx);
        }
      } else if (annexb) {
        H264Index idx;
        ok = index_bytestream(m.data(), m.size(), idx).success();
        if (ok) {
          std::vector<u8> out;
          mux_mp4(m.data(), m.size(), idx, 30, 1, out);
          ok = check_index(idx, m.size()).success();
          if (!ok) {
            fprintf(stderr, "index_bytestream produced an index check_index rejects\n");
            return 1;
          }
        }
      } else {
        tables::VideoDescriptor vd;
        tables::TableDescriptor td;
        tables::DatabaseDescriptor dd;
        td.ParseFromArray(m.data(), (int)m.size());
        dd.ParseFromArray(m.data(), (int)m.size());
        if (vd.ParseFromArray(m.data(), (int)m.size())) {
          H264Index idx;
          ok = index_from_descriptor(vd, idx).success() && check_index(idx, 1 << 20).success();
          std::string again;
          vd.SerializeToString(&again);
        }
      }
      (ok ? accepted : rejected)++;
    }
  }
  printf("accepted %ld rejected %ld\n", accepted, rejected);
  return 0;
}
