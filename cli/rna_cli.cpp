// rna_cli.cpp — command-line front ends with the reference's options and output formats, over the C ABI
// (include/rna_algos_b200.h).  One binary, three programs (chosen by the executable's name or the first argument):
//   mccaskill_algo  -i FASTA -o FILE [-c] [-t N]        reference src/bin/mccaskill_algo.rs:6-113
//   centroid_fold   -i FASTA -o DIR  [-g GAMMA] [-c] [-t N]       src/bin/centroid_fold.rs:13-207
//   durbin_algo     -i FASTA -o FILE [-t N]                        src/bin/durbin_algo.rs:6-90
// Differences that cannot be observed in a reference run: records of a hash map are written in sorted key order
// (the reference iterates hashbrown maps, whose order is unspecified); -t limits the number of GPUs the sequences are
// fanned out over (default: all visible); a non-ACGU base or an unreadable file ends the program with a message and exit status 1 instead
// of a panic.  Score tables are run-time blobs: --tables DIR, else $RNA_ALGOS_B200_TABLES, else ../rna_algos_b200/
// tables_standin next to the executable (written by `python -m rna_algos_b200.tables dump`; --standin-tables only).
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

#include "../include/rna_algos_b200.h"

namespace {

struct Fasta {
  std::vector<std::string> ids;
  std::vector<uint8_t> bases;       // concatenated codes 0..3
  std::vector<uint32_t> offsets;    // n + 1
  uint32_t n() const { return (uint32_t)ids.size(); }
  uint32_t len(uint32_t s) const { return offsets[s + 1] - offsets[s]; }
};

[[noreturn]] void die(const std::string& msg) {
  fprintf(stderr, "error: %s\n", msg.c_str());
  exit(1);
}

// bio::io::fasta semantics as the reference uses them: '>' starts a record, id = first word of the header, the
// sequence is every following line up to the next '>' with line ends removed; bytes2seq (src/utils.rs:562-577)
// accepts a/c/g/u in either case and nothing else.
Fasta read_fasta(const std::string& path) {
  std::ifstream in(path);
  if (!in) die("cannot open " + path);
  Fasta f;
  f.offsets.push_back(0);
  std::string line;
  bool open_rec = false;
  size_t lineno = 0;
  while (std::getline(in, line)) {
    lineno++;
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (!line.empty() && line[0] == '>') {
      if (open_rec) f.offsets.push_back((uint32_t)f.bases.size());
      size_t e = 1;
      while (e < line.size() && line[e] != ' ' && line[e] != '\t') e++;
      f.ids.push_back(line.substr(1, e - 1));
      open_rec = true;
      continue;
    }
    if (line.empty()) continue;
    if (!open_rec) die(path + ":" + std::to_string(lineno) + ": sequence data before the first '>' header");
    for (char ch : line) {
      uint8_t code;
      switch (ch) {
        case 'a': case 'A': code = RNA_BASE_A; break;
        case 'c': case 'C': code = RNA_BASE_C; break;
        case 'g': case 'G': code = RNA_BASE_G; break;
        case 'u': case 'U': code = RNA_BASE_U; break;
        default:
          die(path + ":" + std::to_string(lineno) + ": base '" + std::string(1, ch) + "' is not one of ACGU (the reference panics here, src/utils.rs:570-572)");
      }
      f.bases.push_back(code);
    }
  }
  if (open_rec) f.offsets.push_back((uint32_t)f.bases.size());
  if (f.ids.empty()) die(path + ": no FASTA records");
  for (uint32_t s = 0; s < f.n(); s++) {
    if (f.len(s) == 0) die(path + ": record " + f.ids[s] + " has no sequence");
    if (f.len(s) > RNA_MAX_SEQ_LEN) die(path + ": record " + f.ids[s] + " is longer than 65535 nt");
  }
  return f;
}

// Rust's `{}` for f32: the shortest decimal string that parses back to the same f32, never in exponent form.
// Appends to `out` without temporaries (the output files run to gigabytes).
void append_f32(std::string& out, float x) {
  if (std::isnan(x)) { out += "NaN"; return; }
  if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
  // shortest round-trip digits d.ddd e±xx, laid out positionally (zero padding instead of exact integer digits)
  char buf[48];
  const auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
  const char* p = buf;
  if (*p == '-') { out += '-'; p++; }
  char digits[16];
  int nd = 0;
  const char* e = p;
  for (; e < r.ptr && *e != 'e'; e++) if (*e != '.') digits[nd++] = *e;
  int exp10 = 0;
  std::from_chars(e + 1 + (e[1] == '+' ? 1 : 0), r.ptr, exp10);   // value = d1.d2d3... x 10^exp10
  bool all_zero = true;
  for (int k = 0; k < nd; k++) if (digits[k] != '0') all_zero = false;
  if (all_zero) { out += '0'; return; }
  const int point = exp10 + 1;                                     // digits before the decimal point
  if (point <= 0) { out += "0."; out.append((size_t)(-point), '0'); out.append(digits, (size_t)nd); }
  else if (point >= nd) { out.append(digits, (size_t)nd); out.append((size_t)(point - nd), '0'); }
  else { out.append(digits, (size_t)point); out += '.'; out.append(digits + point, (size_t)(nd - point)); }
}
std::string f32_display(float x) { std::string s; append_f32(s, x); return s; }
void append_uint(std::string& out, uint64_t v) {
  char buf[24];
  const auto r = std::to_chars(buf, buf + sizeof buf, v);
  out.append(buf, (size_t)(r.ptr - buf));
}

struct Opts {
  std::string in, out, tables;
  bool contra = false, has_gamma = false, help = false, standin = false;
  int num_gpus = 0;   // 0 = all visible
  float gamma = 0.f;
};

void usage(const char* prog, bool with_gamma, bool with_model) {
  printf("Usage: %s [options]\n\nOptions:\n", prog);
  printf("    -i, --input_file_path STR   An input FASTA file path containing RNA sequences\n");
  printf("    -o, --output_%s STR  An output %s path\n", with_gamma ? "dir_path " : "file_path", with_gamma ? "directory" : "file");
  if (with_gamma) printf("    -g, --centroid_threshold FLOAT  A specific centroid threshold rather than a range of centroid thresholds\n");
  printf("    -t, --num_threads UINT      The number of GPUs to fan the sequences out over (default: all visible GPUs)\n");
  if (with_model) printf("    -c, --uses_contra_model     Use the CONTRAfold model instead of Turner's model to score RNA secondary structures\n");
  printf("        --tables DIR            Directory with the genuine turner2004.tbl / contrafold_v202.tbl score-table blobs\n");
  printf("                                (written from rna-ss-params by tools/ref_dump; default: $RNA_ALGOS_B200_TABLES)\n");
  printf("        --standin-tables        Run on the in-repo STAND-IN tables instead (NOT the reference's numbers)\n");
  printf("    -h, --help                  Print a help menu\n");
}

Opts parse(int argc, char** argv, int first, bool with_gamma, bool with_model) {
  Opts o;
  for (int k = first; k < argc; k++) {
    const std::string a = argv[k];
    auto val = [&]() -> std::string {
      if (k + 1 >= argc) die("option " + a + " needs a value");
      return argv[++k];
    };
    if (a == "-i" || a == "--input_file_path") o.in = val();
    else if (a == "-o" || a == "--output_file_path" || a == "--output_dir_path") o.out = val();
    else if (a == "-t" || a == "--num_threads") {
      const std::string v = val();
      char* end = nullptr;
      const long t = strtol(v.c_str(), &end, 10);
      if (end == v.c_str() || *end || t < 1) die("cannot parse the number of threads '" + v + "'");
      o.num_gpus = (int)std::min<long>(t, 64);
    }
    else if (with_model && (a == "-c" || a == "--uses_contra_model")) o.contra = true;
    else if (with_gamma && (a == "-g" || a == "--centroid_threshold")) {
      const std::string v = val();
      char* end = nullptr;
      o.gamma = strtof(v.c_str(), &end);
      if (end == v.c_str() || *end) die("cannot parse centroid threshold '" + v + "'");
      o.has_gamma = true;
    } else if (a == "--tables") o.tables = val();
    else if (a == "--standin-tables") o.standin = true;
    else if (a == "-h" || a == "--help") o.help = true;
    else die("unknown option " + a);
  }
  return o;
}

std::string exe_dir() {
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  if (n <= 0) return ".";
  buf[n] = 0;
  std::string p(buf);
  const size_t s = p.rfind('/');
  return s == std::string::npos ? "." : p.substr(0, s);
}

template <class T>
void load_table(const std::string& dir, const char* name, uint32_t kind, T* out) {
  const std::string path = dir + "/" + name;
  std::ifstream in(path, std::ios::binary);
  if (!in) die("cannot open score table " + path + " (genuine blobs: tools/ref_dump; stand-ins: `python -m rna_algos_b200.tables dump DIR`)");
  char magic[8];
  uint32_t hdr[2];
  in.read(magic, 8);
  in.read(reinterpret_cast<char*>(hdr), 8);
  if (!in || memcmp(magic, "RNATBL01", 8) != 0 || hdr[0] != kind || hdr[1] != sizeof(T))
    die(path + ": not a score-table blob of the expected kind and size");
  in.read(reinterpret_cast<char*>(out), sizeof(T));
  if (!in) die(path + ": truncated");
}

// All visible GPUs behind one rna_multi (the reference fans its units out over a thread pool inside the binary,
// src/bin/centroid_fold.rs:104-161; here the library partitions them over the devices).
struct Session {
  rna_multi* h = nullptr;
  void check(int rc, const char* what) {
    if (rc != RNA_OK) die(std::string(what) + ": " + (h ? rna_multi_last_error(h) : "no handle") + " (status " + std::to_string(rc) + ")");
  }
  void open(const Opts& o, bool need_fold_tables) {
    // score tables first (an argument error must not depend on the machine)
    static RnaContraTables ct;
    static RnaTurnerTables tt;
    if (need_fold_tables) {
      std::string dir = o.tables;
      if (dir.empty() && getenv("RNA_ALGOS_B200_TABLES")) dir = getenv("RNA_ALGOS_B200_TABLES");
      // The file names turner2004.tbl / contrafold_v202.tbl are reserved for the genuine rna-ss-params values.  The
      // stand-in tables of this repository give results that differ from the reference's: explicit opt-in only.
      const char* tname = "turner2004.tbl";
      const char* cname = "contrafold_v202.tbl";
      if (o.standin) {
        dir = exe_dir() + "/../rna_algos_b200/tables_standin";
        tname = "standin_turner.tbl";
        cname = "standin_contrafold.tbl";
        fprintf(stderr, "WARNING: --standin-tables: these are NOT the Turner 2004 / CONTRAfold v2.02 parameters of rna-ss-params; "
                        "probabilities and structures differ from the reference's.\n");
      } else if (dir.empty()) {
        die("no score tables: pass --tables DIR (or set RNA_ALGOS_B200_TABLES) with the genuine turner2004.tbl / "
            "contrafold_v202.tbl written by tools/ref_dump, or opt in to the stand-in tables with --standin-tables");
      }
      if (o.contra) load_table(dir, cname, 2, &ct); else load_table(dir, tname, 1, &tt);
    }
    // every visible device, or the first -t / --num_threads of them (a GPU is this implementation's unit of fan-out)
    std::vector<int> devs;
    for (int d = 0; d < o.num_gpus; d++) devs.push_back(d);
    int rc = rna_multi_create(devs.empty() ? nullptr : devs.data(), (int)devs.size(), &h);
    if (rc == RNA_ERR_BAD_ARG && !devs.empty()) rc = rna_multi_create(nullptr, 0, &h);   // (-t beyond the device count)
    if (rc != RNA_OK) die("no usable CUDA device (status " + std::to_string(rc) + "); this program has no CPU path");
    if (need_fold_tables) {
      if (o.contra) check(rna_multi_set_contra_tables(h, &ct), "rna_set_contra_tables");
      else check(rna_multi_set_turner_tables(h, &tt), "rna_set_turner_tables");
    } else {
      RnaAlignTables at;
      rna_align_tables_contralign_v201(&at);
      check(rna_multi_set_align_tables(h, &at), "rna_set_align_tables");
    }
  }
  ~Session() { if (h) rna_multi_destroy(h); }
};

void write_file(const std::string& path, const std::string& data) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) die("cannot write " + path);
  fwrite(data.data(), 1, data.size(), f);
  fclose(f);
}

std::vector<uint64_t> bpp_offsets_of(const Fasta& fa) {
  std::vector<uint64_t> off(fa.n() + 1, 0);
  for (uint32_t s = 0; s < fa.n(); s++) off[s + 1] = off[s] + rna_bpp_len(fa.len(s));
  return off;
}

// ---- mccaskill_algo: src/bin/mccaskill_algo.rs:94-113 --------------------------------------------------------
int main_mccaskill(int argc, char** argv, int first) {
  const Opts o = parse(argc, argv, first, false, true);
  if (o.help) { usage("mccaskill_algo", false, true); return 0; }
  if (o.in.empty() || o.out.empty()) { usage("mccaskill_algo", false, true); die("-i and -o are required"); }
  const Fasta fa = read_fasta(o.in);
  Session S;
  S.open(o, true);
  const std::vector<uint64_t> off = bpp_offsets_of(fa);
  std::vector<float> bpp(off.back());
  S.check(rna_multi_mccaskill_centroid_batch(S.h, fa.bases.data(), fa.offsets.data(), fa.n(), o.contra ? RNA_MODEL_CONTRA : RNA_MODEL_TURNER,
                                             0, nullptr, 0, nullptr, bpp.data(), off.data(), nullptr, nullptr), "rna_multi_mccaskill_centroid_batch");
  std::string buf = "# Format = >{RNA sequence id} {line break} {basepairing left nucleotide}, {basepairing right nucleotide}, {basepairing probability} ...";
  for (uint32_t s = 0; s < fa.n(); s++) {
    buf += "\n\n>" + std::to_string(s) + "\n";
    const uint64_t L = fa.len(s);
    const float* p = bpp.data() + off[s];
    for (uint64_t i = 0; i + 1 < L; i++)
      for (uint64_t j = i + 1; j < L; j++) {
        const float x = p[rna_bpp_index(L, i, j)];
        if (x == RNA_BPP_ABSENT) continue;   // key not in the reference's SparseProbMat (a present 0.0 is written)
        append_uint(buf, i); buf += ','; append_uint(buf, j); buf += ','; append_f32(buf, x); buf += ' ';
      }
  }
  write_file(o.out, buf);
  return 0;
}

// ---- centroid_fold: src/bin/centroid_fold.rs:104-207 ---------------------------------------------------------
int main_centroid(int argc, char** argv, int first) {
  const Opts o = parse(argc, argv, first, true, true);
  if (o.help) { usage("centroid_fold", true, true); return 0; }
  if (o.in.empty() || o.out.empty()) { usage("centroid_fold", true, true); die("-i and -o are required"); }
  const Fasta fa = read_fasta(o.in);
  std::vector<float> gammas;
  if (o.has_gamma && !(std::isinf(o.gamma) && o.gamma < 0)) gammas.push_back(o.gamma);
  else for (int p = -7; p <= 10; p++) gammas.push_back(std::ldexp(1.0f, p));   // (2. as Prob).powi(pow_2), MIN_POW_2..=MAX_POW_2
  Session S;
  S.open(o, true);
  const size_t total = fa.bases.size();
  std::vector<uint8_t> structs(gammas.size() * total);
  S.check(rna_multi_mccaskill_centroid_batch(S.h, fa.bases.data(), fa.offsets.data(), fa.n(), o.contra ? RNA_MODEL_CONTRA : RNA_MODEL_TURNER, 0,
                                       gammas.data(), (uint32_t)gammas.size(), nullptr, nullptr, nullptr, structs.data(), nullptr),
          "rna_mccaskill_centroid_batch");
  struct stat st;
  if (stat(o.out.c_str(), &st) != 0 && mkdir(o.out.c_str(), 0777) != 0) die("cannot create directory " + o.out);
  for (size_t g = 0; g < gammas.size(); g++) {
    std::string buf;
    for (uint32_t s = 0; s < fa.n(); s++) {
      buf += ">" + std::to_string(s) + "\n";
      buf.append(reinterpret_cast<const char*>(structs.data() + g * total + fa.offsets[s]), fa.len(s));
      if (s + 1 < fa.n()) buf += "\n";
    }
    write_file(o.out + "/centroid_threshold=" + f32_display(gammas[g]) + ".fa", buf);
  }
  return 0;
}

// ---- durbin_algo: src/bin/durbin_algo.rs:42-90 ----------------------------------------------------------------
int main_durbin(int argc, char** argv, int first) {
  const Opts o = parse(argc, argv, first, false, false);
  if (o.help) { usage("durbin_algo", false, false); return 0; }
  if (o.in.empty() || o.out.empty()) { usage("durbin_algo", false, false); die("-i and -o are required"); }
  const Fasta fa = read_fasta(o.in);
  Session S;
  S.open(o, false);
  std::vector<uint32_t> pairs;
  std::vector<uint64_t> poff(1, 0);
  for (uint32_t a = 0; a < fa.n(); a++)
    for (uint32_t b = a + 1; b < fa.n(); b++) {
      pairs.push_back(a); pairs.push_back(b);
      poff.push_back(poff.back() + (uint64_t)(fa.len(a) + 2) * (fa.len(b) + 2));
    }
  const uint32_t np = (uint32_t)(pairs.size() / 2);
  std::vector<float> probs(poff.back());
  if (np) S.check(rna_multi_durbin_batch(S.h, fa.bases.data(), fa.offsets.data(), fa.n(), pairs.data(), np, probs.data(), poff.data()), "rna_durbin_batch");
  std::string buf = "# Format = >{RNA sequence id 1},{RNA sequence id 2} {line break} {nucleotide 1}, {nucleotide 2}, {nucletide matching probability} ...";
  for (uint32_t p = 0; p < np; p++) {
    const uint32_t a = pairs[2 * p], b = pairs[2 * p + 1];
    buf += "\n\n>" + std::to_string(a) + "," + std::to_string(b) + "\n";
    const uint64_t n = fa.len(a) + 2, m = fa.len(b) + 2;
    const float* x = probs.data() + poff[p];
    for (uint64_t i = 0; i < n; i++)
      for (uint64_t j = 0; j < m; j++)
        if (x[i * m + j] > 0.f) { append_uint(buf, i - 1); buf += ','; append_uint(buf, j - 1); buf += ','; append_f32(buf, x[i * m + j]); buf += ' '; }
  }
  write_file(o.out, buf);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  std::string prog = argv[0];
  const size_t s = prog.rfind('/');
  if (s != std::string::npos) prog = prog.substr(s + 1);
  int first = 1;
  if (prog != "mccaskill_algo" && prog != "centroid_fold" && prog != "durbin_algo") {
    if (argc < 2) {
      fprintf(stderr, "usage: %s {mccaskill_algo|centroid_fold|durbin_algo} [options]   (%s)\n", argv[0], rna_version());
      return 1;
    }
    prog = argv[1];
    first = 2;
  }
  if (prog == "_parse") {   // test hook: print the records the FASTA reader produces (no GPU needed)
    if (argc <= first) return 1;
    const Fasta fa = read_fasta(argv[first]);
    for (uint32_t q = 0; q < fa.n(); q++) {
      printf("%s %u ", fa.ids[q].c_str(), fa.len(q));
      for (uint32_t x = fa.offsets[q]; x < fa.offsets[q + 1]; x++) putchar("ACGU"[fa.bases[x]]);
      putchar('\n');
    }
    return 0;
  }
  if (prog == "_fmt") {   // test hook: print f32 bit patterns (hex) the way the output files do
    for (int k = first; k < argc; k++) {
      const uint32_t b = (uint32_t)strtoul(argv[k], nullptr, 16);
      float x;
      memcpy(&x, &b, 4);
      printf("%s\n", f32_display(x).c_str());
    }
    return 0;
  }
  if (prog == "mccaskill_algo") return main_mccaskill(argc, argv, first);
  if (prog == "centroid_fold") return main_centroid(argc, argv, first);
  if (prog == "durbin_algo") return main_durbin(argc, argv, first);
  fprintf(stderr, "unknown program %s\n", prog.c_str());
  return 1;
}
