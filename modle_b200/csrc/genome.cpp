// Host input layer of the C ABI (include/modle_b200.h, "genome import"): chrom.sizes, the
// extrusion-barrier BED6 and the optional genomic-intervals BED3 -> intervals with their barriers,
// in the reference's processing order. No device code. Each piece names the reference lines whose
// behaviour it reproduces (paths relative to the reference checkout):
//
//   chrom_sizes::Parser::parse_all              src/libmodle_io/chrom_sizes.cpp:24-66
//   bed::Parser (header skipping, duplicates)   src/libmodle_io/bed.cpp:411-590
//   bed::BED(record, ...) field parsing         src/libmodle_io/bed.cpp:44-330
//   Genome::Genome / import_* / map_barriers_*  src/libmodle/internal/genome.cpp:299-488
//   generate_barriers_from_bed_records          src/libmodle/internal/genome.cpp:423-469
//   override_extrusion_barrier_occupancy        src/libmodle/cpu/simulation.cpp:51-60
//
// Files are read as plain text (the reference also reads gz/bz2/xz/... through libarchive, which
// this image does not have; decompress first).
#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <system_error>
#include <tuple>
#include <vector>

#include "../../include/modle_b200.h"
#include "status.hpp"

namespace {

using modle_b200::fail;
using u64 = std::uint64_t;

struct ParseError {
  std::string msg;
};

// modle::strip_trailing_whitespace: trailing blanks, tabs, CR / LF
std::string_view strip_trailing_whitespace(std::string_view s) {
  while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r' ||
                        s.back() == '\n' || s.back() == '\v' || s.back() == '\f'))
    s.remove_suffix(1);
  return s;
}

// utils::strip_quote_pairs (src/common/utils_impl.hpp:204-214)
std::string_view strip_quote_pairs(std::string_view s) {
  if (s.size() < 2) return s;
  const bool b = s.front() == '\'' || s.front() == '"';
  const bool e = s.back() == '\'' || s.back() == '"';
  return (b && e) ? s.substr(1, s.size() - 2) : s;
}

std::vector<std::string_view> split(std::string_view s, std::string_view seps, bool drop_empty) {
  std::vector<std::string_view> out;
  size_t i = 0;
  for (;;) {
    const size_t j = s.find_first_of(seps, i);
    const std::string_view tok = s.substr(i, j == std::string_view::npos ? j : j - i);
    if (!drop_empty || !tok.empty()) out.push_back(tok);
    if (j == std::string_view::npos) break;
    i = j + 1;
  }
  return out;
}

// utils::parse_numeric_or_throw (src/common/numeric_utils_impl.hpp:75-81): from_chars, and an
// exception only when the token was not consumed to its end AND from_chars reported an error
// (so "12abc" is 12, and an out-of-range literal leaves the field untouched).
template <class N>
void parse_numeric_or_throw(std::string_view tok, N& field) {
  const char* first = tok.data();
  const char* last = tok.data() + tok.size();
  const auto res = std::from_chars(first, last, field);
  if (res.ptr != last && res.ec != std::errc{})
    throw ParseError{"Unable to convert field \"" + std::string(tok) + "\" to a number"};
}

struct BedRecord {
  std::string chrom;
  u64 start = 0, end = 0;
  std::string name;
  double score = 0.0;
  char strand = '.';
  size_t line = 0;
};

// bed_strand_encoding (src/libmodle_io/include/bed/modle/bed/bed.hpp:251-264)
char parse_strand_or_throw(std::string_view tok) {
  tok = strip_quote_pairs(tok);
  static const std::map<std::string_view, char> enc = {
      {"+", '+'},       {"plus", '+'},    {"fwd", '+'},     {"Fwd", '+'},     {"forward", '+'},
      {"Forward", '+'}, {"FWD", '+'},     {"FORWARD", '+'}, {"-", '-'},       {"minus", '-'},
      {"rev", '-'},     {"Rev", '-'},     {"reverse", '-'}, {"Reverse", '-'}, {"REV", '-'},
      {"REVERSE", '-'}, {".", '.'},       {"", '.'},        {"none", '.'},    {"None", '.'},
      {"NONE", '.'},    {"unknown", '.'}, {"Unknown", '.'}, {"unk", '.'},     {"Unk", '.'},
      {"UNK", '.'}};
  const auto it = enc.find(tok);
  if (it == enc.end()) throw ParseError{"unrecognized strand \"" + std::string(tok) + "\""};
  return it->second;
}

// BED::BED(record, id, standard, validate) restricted to the BED3 / BED6 dialects the simulation
// asks for (genome.cpp:312, :392). `dialect` is 3 or 6.
BedRecord parse_bed_record(std::string_view line, int dialect) {
  const std::vector<std::string_view> toks =
      split(strip_trailing_whitespace(line), "\t ", /*drop_empty=*/true);
  // validate_record: detect_standard throws below 3 fields; field counts that are not a BED
  // standard (7, 8, 10, 11, > 12) count as "none" = 254 and pass the "at least" test
  if (toks.size() < 3)
    throw ParseError{"expected at least 3 fields, got " + std::to_string(toks.size())};
  const size_t n = toks.size();
  const size_t detected = (n == 3 || n == 4 || n == 5 || n == 6 || n == 9 || n == 12) ? n : 254;
  if (detected < static_cast<size_t>(dialect))
    throw ParseError{"Invalid BED record detected: Expected BED record with at least " +
                     std::to_string(dialect) + " fields, got " + std::to_string(n)};
  BedRecord r;
  try {
    r.chrom = std::string(strip_quote_pairs(toks[0]));
    parse_numeric_or_throw(toks[1], r.start);
    parse_numeric_or_throw(toks[2], r.end);
    if (r.start > r.end)
      throw ParseError{"Invalid BED record detected: chrom_start > chrom_end: chrom=\"" + r.chrom +
                       "\"; start=" + std::to_string(r.start) + "; end=" + std::to_string(r.end)};
    if (dialect == 3) return r;
    r.name = std::string(strip_quote_pairs(toks[3]));
    parse_numeric_or_throw(toks[4], r.score);
    if (r.score < 0 || r.score > 1000)
      throw ParseError{
          "Invalid BED record detected: score field should be between 0.0 and 1000.0, is " +
          std::to_string(r.score) + "."};
    r.strand = parse_strand_or_throw(toks[5]);
  } catch (const ParseError& e) {
    throw ParseError{"An error occurred while parsing the following BED record \"" +
                     std::string(line) + "\":\n  " + e.msg};
  }
  return r;
}

std::vector<std::string> read_lines(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw ParseError{"Unable to open file " + path + " for reading"};
  std::vector<std::string> lines;
  std::string buff;
  while (std::getline(f, buff)) lines.push_back(buff);
  return lines;
}

// bed::Parser: skip_header (leading empty lines, '#...', lines containing "track" or "browser"),
// then one record per non-empty line; a second record with the same (chrom, start, end) is an
// error (BED::operator< compares exactly those, bed.cpp:354-366).
std::vector<BedRecord> parse_bed_file(const std::string& path, int dialect) {
  const std::vector<std::string> lines = read_lines(path);
  size_t i = 0;
  for (; i < lines.size(); ++i) {
    const std::string& l = lines[i];
    if (l.empty()) continue;
    if (l.front() == '#' || l.find("track") != std::string::npos ||
        l.find("browser") != std::string::npos)
      continue;
    break;
  }
  std::vector<BedRecord> out;
  std::map<std::tuple<std::string, u64, u64>, size_t> seen;
  for (; i < lines.size(); ++i) {
    if (lines[i].empty()) continue;
    BedRecord r;
    try {
      r = parse_bed_record(lines[i], dialect);
    } catch (const ParseError& e) {
      throw ParseError{e.msg + " (line " + std::to_string(i + 1) + " of file " + path + ")"};
    }
    r.line = i + 1;
    const auto [it, fresh] = seen.emplace(std::make_tuple(r.chrom, r.start, r.end), r.line);
    if (!fresh)
      throw ParseError{"Detected duplicate record at line " + std::to_string(r.line) + " of file " +
                       path + ". First occurrence was at line " + std::to_string(it->second) + "."};
    out.push_back(std::move(r));
  }
  return out;
}

struct Chromosome {
  std::string name;
  u64 size;
};

// chrom_sizes::Parser::parse_all (chrom_sizes.cpp:24-66): tab separated, exactly two fields,
// quotes stripped from the name, duplicates and zero lengths rejected; the size goes through the
// BED3 parser as "name\t0\tsize".
std::vector<Chromosome> parse_chrom_sizes(const std::string& path) {
  std::vector<Chromosome> out;
  std::map<std::string, size_t> seen;
  const std::vector<std::string> lines = read_lines(path);
  for (size_t i = 0; i < lines.size(); ++i) {
    const std::string_view buff = strip_trailing_whitespace(lines[i]);
    if (buff.empty()) continue;
    try {
      const auto toks = split(buff, "\t", /*drop_empty=*/false);
      if (toks.size() != 2)
        throw ParseError{"expected exactly 2 fields, found " + std::to_string(toks.size()) +
                         ": \"" + std::string(buff) + "\""};
      const std::string name(strip_quote_pairs(toks[0]));
      if (seen.count(name))
        throw ParseError{"found multiple records for chrom \"" + name + "\""};
      if (toks[1] == "0") throw ParseError{"chrom \"" + name + "\" has a length of 0bp"};
      const BedRecord r = parse_bed_record(name + "\t0\t" + std::string(toks[1]), 3);
      seen.emplace(name, i + 1);
      out.push_back(Chromosome{r.chrom, r.end});
    } catch (const ParseError& e) {
      throw ParseError{"encountered a malformed record at line " + std::to_string(i + 1) +
                       " of file " + path + ": " + e.msg + ".\n Line that triggered the error:\n\"" +
                       std::string(buff) + "\""};
    }
  }
  if (out.empty()) throw ParseError{"Unable to import any chromosome from " + path + "!"};
  return out;
}

struct Interval {
  size_t chrom_id;
  u64 start, end;
  std::vector<modle_b200_barrier> barriers;
};

// BED_tree<>::find_overlaps(chrom, start, end) (bed_impl.hpp:166-180 over IITree::equal_range,
// src/interval_tree/interval_tree_impl.hpp:150-215): the chromosome's records ordered by start;
// the result is the CONTIGUOUS span from the first to the last record with
// record.start < end && start < record.end -- records in between that do not overlap the query
// themselves (they end at or before `start`) are part of the span, as in the reference.
std::vector<const BedRecord*> find_overlaps(const std::vector<BedRecord>& records,
                                            const std::string& chrom, u64 start, u64 end) {
  std::vector<const BedRecord*> sorted;
  for (const auto& r : records)
    if (r.chrom == chrom) sorted.push_back(&r);
  std::stable_sort(sorted.begin(), sorted.end(), [](const BedRecord* a, const BedRecord* b) {
    return a->start != b->start ? a->start < b->start : a->end < b->end;
  });
  size_t first = sorted.size(), last = 0;
  for (size_t i = 0; i < sorted.size(); ++i) {
    if (sorted[i]->start < end && start < sorted[i]->end) {
      first = std::min(first, i);
      last = i + 1;
    }
  }
  if (first == sorted.size()) return {};
  return std::vector<const BedRecord*>(sorted.begin() + static_cast<std::ptrdiff_t>(first),
                                       sorted.begin() + static_cast<std::ptrdiff_t>(last));
}

}  // namespace

struct modle_b200_genome {
  std::vector<Chromosome> chroms;
  std::vector<Interval> intervals;
  std::vector<u64> chrom_first_bin;  // cooler bin table: ceil(size / bin_size) bins per chromosome
  u64 bin_size = 0;
  u64 num_barriers = 0;
};

extern "C" {

int modle_b200_genome_import(const char* path_to_chrom_sizes, const char* path_to_extr_barriers,
                             const char* path_to_genomic_intervals,
                             const modle_b200_sim_params* params,
                             int interpret_name_field_as_puu, modle_b200_genome** out) {
  if (!out) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!path_to_chrom_sizes || !path_to_extr_barriers || !params)
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "NULL argument");
  if (params->bin_size == 0) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "bin_size is 0");
  try {
    auto g = std::make_unique<modle_b200_genome>();
    g->bin_size = params->bin_size;
    g->chroms = parse_chrom_sizes(path_to_chrom_sizes);
    u64 bin = 0;
    for (const auto& c : g->chroms) {
      g->chrom_first_bin.push_back(bin);
      bin += (c.size + params->bin_size - 1) / params->bin_size;
    }
    // Genome::import_genomic_intervals (genome.cpp:363-421): whole chromosomes without a BED;
    // otherwise, chromosome by chromosome, the records overlapping [0, size) in tree order
    if (!path_to_genomic_intervals || !*path_to_genomic_intervals) {
      for (size_t c = 0; c < g->chroms.size(); ++c)
        g->intervals.push_back(Interval{c, 0, g->chroms[c].size, {}});
    } else {
      const auto records = parse_bed_file(path_to_genomic_intervals, 3);
      for (size_t c = 0; c < g->chroms.size(); ++c)
        for (const BedRecord* r : find_overlaps(records, g->chroms[c].name, 0, g->chroms[c].size))
          g->intervals.push_back(Interval{c, r->start, r->end, {}});
      if (g->intervals.empty())
        throw ParseError{std::string("unable to import any interval from ") +
                         path_to_genomic_intervals + "!"};
    }
    // Genome::map_barriers_to_intervals + generate_barriers_from_bed_records (genome.cpp:423-488)
    const auto barriers = parse_bed_file(path_to_extr_barriers, 6);
    const double pbb = params->barrier_occupied_stp, puu = params->barrier_not_occupied_stp;
    for (auto& iv : g->intervals) {
      for (const BedRecord* r :
           find_overlaps(barriers, g->chroms[iv.chrom_id].name, iv.start, iv.end)) {
        const std::string where = "found invalid extrusion barrier " + r->chrom + "\t" +
                                  std::to_string(r->start) + "\t" + std::to_string(r->end) + ": ";
        if (r->strand == '.') continue;
        if (r->score < 0 || r->score > 1)
          throw ParseError{where + "invalid score field: expected a score between 0 and 1"};
        if (interpret_name_field_as_puu) {
          // the reference validates the name as a number in [0, 1] and then does not use it
          // (compute_barrier_stp receives the default, genome.cpp:456-457)
          double v = puu;
          bool ok = true;
          try {
            parse_numeric_or_throw(std::string_view(r->name), v);
          } catch (const ParseError&) {
            ok = false;
          }
          if (!ok || v < 0 || v > 1)
            throw ParseError{where +
                             "invalid name field: expected name to be a number between 0 and 1, "
                             "found " + r->name + "."};
        }
        modle_b200_barrier b{};
        b.pos = (r->start + r->end + 1) / 2;
        // compute_barrier_stp (genome.cpp:255-271); a '+' motif blocks REV-moving units
        if (r->score != 0.0) {
          b.stp_active = modle_b200_stp_active_from_occupancy(puu, r->score);
          b.stp_inactive = puu;
        } else {
          b.stp_active = pbb;
          b.stp_inactive = puu;
        }
        if (params->override_extrusion_barrier_occupancy) {  // simulation.cpp:51-60
          b.stp_active = pbb;
          b.stp_inactive = puu;
        }
        b.blocking_direction = r->strand == '+' ? MODLE_B200_DIR_REV : MODLE_B200_DIR_FWD;
        iv.barriers.push_back(b);
      }
      // the simulation wants barriers ordered by position (ExtrusionBarriers::sort,
      // extrusion_barriers.cpp:232-257); equal positions keep the tree order
      std::stable_sort(iv.barriers.begin(), iv.barriers.end(),
                       [](const modle_b200_barrier& a, const modle_b200_barrier& b) {
                         return a.pos < b.pos;
                       });
      g->num_barriers += iv.barriers.size();
    }
    *out = g.release();
    return MODLE_B200_OK;
  } catch (const ParseError& e) {
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, e.msg);
  } catch (const std::exception& e) {
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, e.what());
  }
}

void modle_b200_genome_free(modle_b200_genome* g) { delete g; }

size_t modle_b200_genome_num_chromosomes(const modle_b200_genome* g) {
  return g ? g->chroms.size() : 0;
}
size_t modle_b200_genome_num_intervals(const modle_b200_genome* g) {
  return g ? g->intervals.size() : 0;
}
uint64_t modle_b200_genome_num_barriers(const modle_b200_genome* g) {
  return g ? g->num_barriers : 0;
}

int modle_b200_genome_get_interval(const modle_b200_genome* g, size_t i, const char** chrom_name,
                                   size_t* chrom_id, uint64_t* chrom_size, uint64_t* start,
                                   uint64_t* end, const modle_b200_barrier** barriers,
                                   size_t* num_barriers, uint64_t* bin_offset) {
  if (!g) return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "genome is NULL");
  if (i >= g->intervals.size())
    return fail(MODLE_B200_ERR_INVALID_ARGUMENT, "interval index out of range");
  const Interval& iv = g->intervals[i];
  if (chrom_name) *chrom_name = g->chroms[iv.chrom_id].name.c_str();
  if (chrom_id) *chrom_id = iv.chrom_id;
  if (chrom_size) *chrom_size = g->chroms[iv.chrom_id].size;
  if (start) *start = iv.start;
  if (end) *end = iv.end;
  if (barriers) *barriers = iv.barriers.empty() ? nullptr : iv.barriers.data();
  if (num_barriers) *num_barriers = iv.barriers.size();
  // emplace_pixel's offset (contact_matrix_dense_io_impl.hpp:30-43): first bin of the chromosome
  // in the cooler's bin table + interval start / resolution
  if (bin_offset) *bin_offset = g->chrom_first_bin[iv.chrom_id] + iv.start / g->bin_size;
  return MODLE_B200_OK;
}

}  // extern "C"
