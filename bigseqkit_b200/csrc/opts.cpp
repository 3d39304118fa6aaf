// opts.cpp -- option parsing + the Before() validation of the reference operators.
#include "opts.h"

#include <strings.h>

#include <cstdio>
#include <cstring>

#include "bsk.h"
#include "gcode_tables.h"
#include "json.h"

namespace bsk {

// ------------------------------------------------------------------ alphabets
// Restated from the published shenwei356/bio v0.7.0 seq/alphabet.go (not vendored
// in the reference tree; call sites bigseqkit-lib/helper.go:288,305,320).
namespace {
struct AlphabetTables {
  uint8_t valid[AB_COUNT][256];
  uint8_t pair[AB_COUNT][256];
  AlphabetTables() {
    memset(valid, 0, sizeof valid);
    for (int a = 0; a < AB_COUNT; a++)
      for (int i = 0; i < 256; i++) pair[a][i] = (uint8_t)i;
    def(AB_DNA, "acgtACGT", "tgcaTGCA", " -.", "nN.");
    def(AB_DNARED, "acgtryswkmbdhvACGTRYSWKMBDHV", "tgcayrswmkvhdbTGCAYRSWMKVHDB", " -.", "nN.");
    def(AB_RNA, "acguACGU", "ugcaUGCA", " -.", "nN.");
    def(AB_RNARED, "acguryswkmbdhvACGURYSWKMBDHV", "ugcayrswmkvhdbUGCAYRSWMKVHDB", " -.", "nN.");
    const char *prot = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ*_.";
    def(AB_PROTEIN, prot, prot, " -", "xXbBzZ");
    for (int i = 0; i < 256; i++) valid[AB_UNLIMIT][i] = 1;
  }
  void def(int a, const char *letters, const char *pairs, const char *gap, const char *amb) {
    for (size_t i = 0; letters[i]; i++) {
      valid[a][(uint8_t)letters[i]] = 1;
      pair[a][(uint8_t)letters[i]] = (uint8_t)pairs[i];
    }
    for (size_t i = 0; gap[i]; i++) valid[a][(uint8_t)gap[i]] = 1;
    for (size_t i = 0; amb[i]; i++) valid[a][(uint8_t)amb[i]] = 1;
  }
};
const AlphabetTables &tables() {
  static AlphabetTables t;
  return t;
}
}  // namespace

const char *alphabet_name(int a) {
  static const char *n[] = {"", "DNA", "DNAredundant", "RNA", "RNAredundant", "Protein", "Unlimit"};
  return (a >= 0 && a < AB_COUNT) ? n[a] : "";
}
const uint8_t *alphabet_valid(int a) { return tables().valid[a]; }
const uint8_t *alphabet_pair(int a) { return tables().pair[a]; }
void alphabet_class_masks(uint8_t out[256]) {
  static const int order[5] = {AB_DNA, AB_RNA, AB_DNARED, AB_RNARED, AB_PROTEIN};
  for (int c = 0; c < 256; c++) {
    uint8_t m = 0;
    for (int k = 0; k < 5; k++)
      if (tables().valid[order[k]][c]) m |= (uint8_t)(1u << k);
    out[c] = m;
  }
}
int alphabet_from_mask(unsigned m, bool empty) {
  if (empty) return AB_UNLIMIT;
  if (m & 1) return AB_DNARED;   // DNA -> DNAredundant
  if (m & 2) return AB_RNARED;   // RNA -> RNAredundant
  if (m & 4) return AB_DNARED;
  if (m & 8) return AB_RNARED;
  if (m & 16) return AB_PROTEIN;
  return AB_UNLIMIT;
}
static bool all_valid(int a, const std::string &s) {
  for (unsigned char c : s)
    if (!tables().valid[a][c]) return false;
  return true;
}
bool pattern_is_legal(const std::string &s) {
  return all_valid(AB_DNARED, s) || all_valid(AB_RNARED, s) || all_valid(AB_PROTEIN, s);
}

Op op_from_name(const char *n) {
  if (!n) return OP_INVALID;
  if (!strcmp(n, "SeqTransform") || !strcmp(n, "seq")) return OP_SEQ;
  if (!strcmp(n, "Stats") || !strcmp(n, "stats")) return OP_STATS;
  if (!strcmp(n, "RmDup") || !strcmp(n, "rmdup")) return OP_RMDUP;
  if (!strcmp(n, "RmDupPrepare")) return OP_RMDUP_PREPARE;
  if (!strcmp(n, "Translate") || !strcmp(n, "translate")) return OP_TRANSLATE;
  if (!strcmp(n, "Locate") || !strcmp(n, "locate")) return OP_LOCATE;
  if (!strcmp(n, "Grep") || !strcmp(n, "grep")) return OP_GREP;
  if (!strcmp(n, "SubseqTransform") || !strcmp(n, "subseq")) return OP_SUBSEQ;
  if (!strcmp(n, "Fq2Fa") || !strcmp(n, "fq2fa")) return OP_FQ2FA;
  if (!strcmp(n, "Duplicate") || !strcmp(n, "duplicate")) return OP_DUPLICATE;
  if (!strcmp(n, "RangePrepare") || !strcmp(n, "Range") || !strcmp(n, "range") || !strcmp(n, "Head") || !strcmp(n, "head")) return OP_RANGE;
  return OP_INVALID;
}

// ------------------------------------------------------------------ JSON -> Opts
namespace {
void jbool(const JValue &o, const char *k, bool &dst) {
  const JValue *v = o.get(k);
  if (!v) return;
  if (v->kind == JValue::Bool) dst = v->b;
  else if (v->kind == JValue::Num) dst = v->num != 0;
}
void jint(const JValue &o, const char *k, int &dst) {
  const JValue *v = o.get(k);
  if (v && v->kind == JValue::Num) dst = (int)v->num;
}
void ji64(const JValue &o, const char *k, int64_t &dst) {
  const JValue *v = o.get(k);
  if (v && v->kind == JValue::Num) dst = (int64_t)v->num;
}
void jdbl(const JValue &o, const char *k, double &dst) {
  const JValue *v = o.get(k);
  if (v && v->kind == JValue::Num) dst = v->num;
}
void jstr(const JValue &o, const char *k, std::string &dst) {
  const JValue *v = o.get(k);
  if (v && v->kind == JValue::Str) dst = v->str;
}
bool jstrs(const JValue &o, const char *k, std::vector<std::string> &dst) {
  const JValue *v = o.get(k);
  if (!v) return false;
  if (v->kind == JValue::Arr) {
    dst.clear();
    for (auto &e : v->arr) {
      if (e.kind == JValue::Str) dst.push_back(e.str);
      else if (e.kind == JValue::Num) { char t[32]; snprintf(t, sizeof t, "%d", (int)e.num); dst.push_back(t); }
    }
    return true;
  }
  if (v->kind == JValue::Str) {  // tolerate "1,2,3"
    dst.clear();
    size_t s = 0;
    for (;;) {
      size_t c = v->str.find(',', s);
      dst.push_back(v->str.substr(s, c == std::string::npos ? c : c - s));
      if (c == std::string::npos) break;
      s = c + 1;
    }
    return true;
  }
  return false;
}

// KitConfig.GetAlphabet (bigseqkit/helper.go:68-84)
bool seqtype_alphabet(const std::string &t, int &ab, std::string &err) {
  const char *s = t.c_str();
  if (!strcasecmp(s, "auto")) ab = AB_NIL;
  else if (!strcasecmp(s, "dna")) ab = AB_DNARED;
  else if (!strcasecmp(s, "rna")) ab = AB_RNARED;
  else if (!strcasecmp(s, "protein")) ab = AB_PROTEIN;
  else if (!strcasecmp(s, "unlimit")) ab = AB_UNLIMIT;
  else {
    err = "invalid sequence type: " + t + ", available value: dna|rna|protein|unlimit|auto";
    return false;
  }
  return true;
}

// region "a:b" (bigseqkit-lib/subseq.go:78-96, grep.go:93-118)
bool parse_region(const std::string &region, const char *cmd, int &start, int &end, std::string &err) {
  const char *c = region.c_str();
  bool ok = true;
  if (*c == '-') c++;
  if (!(*c >= '0' && *c <= '9')) ok = false;
  while (*c >= '0' && *c <= '9') c++;
  if (*c != ':') ok = false;
  else c++;
  if (*c == '-') c++;
  if (!(*c >= '0' && *c <= '9')) ok = false;
  while (*c >= '0' && *c <= '9') c++;
  if (*c) ok = false;
  if (!ok) {
    err = "invalid region: " + region + ". type \"seqkit " + cmd + " -h\" for more examples";
    return false;
  }
  start = atoi(region.c_str());
  end = atoi(strchr(region.c_str(), ':') + 1);
  if (start == 0 || end == 0) { err = "both start and end should not be 0"; return false; }
  if (start < 0 && end > 0) { err = "when start < 0, end should not > 0"; return false; }
  return true;
}

bool gap_letters_ok(const std::string &g, std::string &err) {
  if (g.empty()) { err = "value of flag -G (--gap-letters) should not be empty"; return false; }
  for (unsigned char c : g)
    if (c > 127) { err = "value of -G (--gap-letters) contains non-ASCII characters"; return false; }
  return true;
}
}  // namespace

bool parse_and_validate(Op op, const char *json, Opts &o, std::string &err, int &code) {
  code = BSK_ERR_ARG;
  JValue root;
  std::string js = (json && *json) ? json : "{}";
  JParser p(js);
  if (!p.parse(root, err)) return false;
  if (root.kind != JValue::Obj) { err = "options must be a JSON object"; return false; }
  o.GapLetters = (op == OP_STATS) ? "- ." : "- \t.";

  if (const JValue *cfg = root.get("Config")) {
    jstr(*cfg, "SeqType", o.SeqType);
    jint(*cfg, "LineWidth", o.LineWidth);
    jstr(*cfg, "IDRegexp", o.IDRegexp);
    jbool(*cfg, "IDNCBI", o.IDNCBI);
    jint(*cfg, "AlphabetGuessSeqLength", o.AlphabetGuessSeqLength);
    jint(*cfg, "ValidateSeqLength", o.ValidateSeqLength);
  }
  // flat keys (option struct fields)
  jbool(root, "Reverse", o.Reverse); jbool(root, "Complement", o.Complement); jbool(root, "Name", o.Name);
  jbool(root, "Seq", o.Seq); jbool(root, "Qual", o.Qual); jbool(root, "OnlyId", o.OnlyId);
  jbool(root, "RemoveGaps", o.RemoveGaps); jstr(root, "GapLetters", o.GapLetters);
  jbool(root, "LowerCase", o.LowerCase); jbool(root, "UpperCase", o.UpperCase);
  jbool(root, "Dna2rna", o.Dna2rna); jbool(root, "Rna2dna", o.Rna2dna); jbool(root, "ValidateSeq", o.ValidateSeq);
  jint(root, "ValidateSeqLength", o.ValidateSeqLength); jint(root, "MaxLen", o.MaxLen); jint(root, "MinLen", o.MinLen);
  jint(root, "QualAsciiBase", o.QualAsciiBase); jdbl(root, "MinQual", o.MinQual); jdbl(root, "MaxQual", o.MaxQual);
  jbool(root, "Tabular", o.Tabular); jbool(root, "All", o.All); jstr(root, "FqEncoding", o.FqEncoding);
  jbool(root, "ByName", o.ByName); jbool(root, "BySeq", o.BySeq); jbool(root, "IgnoreCase", o.IgnoreCase);
  jbool(root, "OnlyPositiveStrand", o.OnlyPositiveStrand);
  jstr(root, "DupSeqsFile", o.DupSeqsFile); jstr(root, "DupNumFile", o.DupNumFile);
  jint(root, "TranslTable", o.TranslTable); jstrs(root, "Frame", o.Frame);
  jbool(root, "Trim", o.Trim); jbool(root, "Clean", o.Clean); jbool(root, "AllowUnknownCodon", o.AllowUnknownCodon);
  jbool(root, "InitCodonAsM", o.InitCodonAsM); jbool(root, "AppendFrame", o.AppendFrame);
  jint(root, "ListTranslTable", o.ListTranslTable);
  jint(root, "ListTranslTableWithAmbCodons", o.ListTranslTableWithAmbCodons);
  jstr(root, "PatternFile", o.PatternFile);
  jbool(root, "Degenerate", o.Degenerate); jbool(root, "UseRegexp", o.UseRegexp); jbool(root, "UseFmi", o.UseFmi);
  jbool(root, "NonGreedy", o.NonGreedy); jbool(root, "HideMatched", o.HideMatched); jbool(root, "Circular", o.Circular);
  jbool(root, "InvertMatch", o.InvertMatch); jbool(root, "Count", o.Count); jbool(root, "DeleteMatched", o.DeleteMatched);
  jint(root, "MaxMismatch", o.MaxMismatch);
  jstr(root, "Region", o.Region);
  // Duplicate: vars "times" (bigseqkit-lib/duplicate.go:20); RangePrepare: vars "start" / "end" (range.go:20-24);
  // Head: N (bigseqkit/head.go:33-44 = Range "1:N", i.e. start 0, end N); IndexBase = global index of the partition's
  // first record (what MapWithIndex supplies)
  ji64(root, "Times", o.Times); ji64(root, "times", o.Times);
  ji64(root, "Start", o.RangeStart); ji64(root, "start", o.RangeStart);
  ji64(root, "End", o.RangeEnd); ji64(root, "end", o.RangeEnd);
  ji64(root, "IndexBase", o.IndexBase);
  if (const JValue *nv = root.get("N")) {
    if (nv->kind == JValue::Num) { o.RangeStart = 0; o.RangeEnd = (int64_t)nv->num; }
  }
  if (op == OP_SUBSEQ) { jstr(root, "Gtf", o.SubseqGtf); jstr(root, "Bed", o.SubseqBed); }
  else { jbool(root, "Gtf", o.Gtf); jbool(root, "Bed", o.Bed); }
  if (o.IDNCBI) o.IDRegexp = "\\|([^\\|]+)\\| ";  // bigseqkit/helper.go:97-100

  // patterns: ["ACGT", ...] (name = pattern, bigseqkit-lib/locate.go:141) or [["name","ACGT"], ...]
  if (const JValue *pv = root.get("Pattern")) {
    if (pv->kind == JValue::Arr)
      for (auto &e : pv->arr) {
        if (e.kind == JValue::Str) { o.PatternNames.push_back(e.str); o.Patterns.push_back(e.str); }
        else if (e.kind == JValue::Arr && e.arr.size() == 2 && e.arr[0].kind == JValue::Str && e.arr[1].kind == JValue::Str) {
          o.PatternNames.push_back(e.arr[0].str);
          o.Patterns.push_back(e.arr[1].str);
        }
      }
  }
  // default Pattern is [""] (bigseqkit/locate.go:30, grep.go:34): a lone empty pattern means "none given"
  if (o.Patterns.size() == 1 && o.Patterns[0].empty()) { o.Patterns.clear(); o.PatternNames.clear(); }

  if (!seqtype_alphabet(o.SeqType, o.alphabet, err)) return false;
  if (o.IDRegexp != "^(\\S+)\\s?" && !o.IDNCBI) {
    code = BSK_ERR_UNSUPPORTED;
    err = "custom --id-regexp is outside the accelerated path (only the default and --id-ncbi are supported)";
    return false;
  }

  switch (op) {
    case OP_SEQ:  // bigseqkit-lib/seq.go:36-76
      if (!gap_letters_ok(o.GapLetters, err)) return false;
      if (o.MinLen >= 0 && o.MaxLen >= 0 && o.MinLen > o.MaxLen) {
        err = "value of flag -m (--min-len) should be >= value of flag -M (--max-len)";
        return false;
      }
      if (o.MinQual >= 0 && o.MaxQual >= 0 && o.MinQual > o.MaxQual) {
        err = "value of flag -Q (--min-qual) should be <= value of flag -R (--max-qual)";
        return false;
      }
      if (o.LowerCase && o.UpperCase) {
        err = "could not give both flags -l (--lower-case) and -u (--upper-case)";
        return false;
      }
      break;
    case OP_STATS: {  // bigseqkit-lib/stats.go:27-46, helper.go:119-136
      if (!gap_letters_ok(o.GapLetters, err)) return false;
      const char *e = o.FqEncoding.c_str();
      if (o.FqEncoding.empty()) o.fq_offset = 0;
      else if (!strcasecmp(e, "sanger") || !strcasecmp(e, "illumina-1.8+")) o.fq_offset = 33;
      else if (!strcasecmp(e, "solexa") || !strcasecmp(e, "illumina-1.3+") || !strcasecmp(e, "illumina-1.5+")) o.fq_offset = 64;
      else {
        err = "unsupported quality encoding: " + o.FqEncoding +
              ". available values: 'sanger', 'solexa', 'illumina-1.3+', 'illumina-1.5+', 'illumina-1.8+'";
        return false;
      }
      break;
    }
    case OP_RMDUP:
    case OP_RMDUP_PREPARE:  // bigseqkit/rmdup.go:79-85
      if (o.BySeq && o.ByName) { err = "only one/none of the flags -s (--by-seq) and -n (--by-name) is allowed"; return false; }
      if (o.OnlyPositiveStrand && !o.BySeq) { err = "flag -s (--by-seq) needed when using -P (--only-positive-strand)"; return false; }
      break;
    case OP_TRANSLATE: {  // bigseqkit-lib/translate.go:43-61
      if (!find_gcode(o.TranslTable)) {
        char t[64];
        snprintf(t, sizeof t, "invalid translate table: %d", o.TranslTable);
        err = t;
        return false;
      }
      if (o.ListTranslTable >= 0 || o.ListTranslTableWithAmbCodons >= 0) {
        code = BSK_ERR_UNSUPPORTED;
        err = "-l/-L (list translate table) is a host-only listing, not part of the accelerated path";
        return false;
      }
      o.frames.clear();
      for (auto &tok : o.Frame) {
        char *endp = nullptr;
        long f = strtol(tok.c_str(), &endp, 10);
        if (tok.empty() || *endp) {
          err = "invalid frame(s): " + tok +
                ". available: 1, 2, 3, -1, -2, -3, and 6 for all. multiple frames should be separated by comma";
          return false;
        }
        if (!(f == 1 || f == 2 || f == 3 || f == -1 || f == -2 || f == -3 || f == 6)) {
          char t[96];
          snprintf(t, sizeof t, "invalid frame: %ld. available: 1, 2, 3, -1, -2, -3, and 6 for all", f);
          err = t;
          return false;
        }
        if (f == 6) { o.frames = {1, 2, 3, -1, -2, -3}; break; }
        o.frames.push_back((int)f);
      }
      if (o.frames.size() > 66) {  // the device-side frame list holds 66 entries (the reference takes any number of repeats)
        code = BSK_ERR_UNSUPPORTED;
        err = "translate: more than 66 frame values are outside the accelerated path";
        return false;
      }
      break;
    }
    case OP_LOCATE: {  // bigseqkit-lib/locate.go:33-193
      if (o.UseRegexp || o.Degenerate || o.MaxMismatch > 0 || o.UseFmi) {
        code = BSK_ERR_UNSUPPORTED;
        err = "locate -r/-d/-m/-F (regexp, degenerate, mismatch, FM-index) are outside the accelerated exact-match path";
        return false;
      }
      if (o.Patterns.empty()) { err = "one of flags -p (--pattern) and -f (--pattern-file) needed"; return false; }
      for (size_t i = 0; i < o.Patterns.size(); i++) {
        std::string &pt = o.Patterns[i];
        if (pt.empty()) { err = "one of flags -p (--pattern) and -f (--pattern-file) needed"; return false; }
        if (o.IgnoreCase)
          for (auto &c : pt)
            if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
        if (pt.find('.') != std::string::npos || !pattern_is_legal(pt)) {
          err = "illegal DNA/RNA/Protein sequence: " + o.PatternNames[i] +
                ", you may switch on -d/--degenerate or -r/--use-regexp";
          return false;
        }
      }
      break;
    }
    case OP_GREP: {  // bigseqkit-lib/grep.go:41-253
      if (o.UseRegexp || o.Degenerate || o.MaxMismatch > 0 || o.DeleteMatched) {
        code = BSK_ERR_UNSUPPORTED;
        err = "grep -r/-d/-m/--delete-matched are outside the accelerated exact-match path";
        return false;
      }
      if (o.Patterns.empty()) { err = "one of flags -p (--pattern) and -f (--pattern-file) needed"; return false; }
      if (!o.Region.empty()) {
        o.has_region = true;
        o.BySeq = true;
        if (!parse_region(o.Region, "grep", o.region_start, o.region_end, err)) return false;
      }
      std::vector<std::string> kept;
      for (auto &pt : o.Patterns) {
        if (pt.empty()) continue;  // pattern file: empty lines skipped (grep.go:136-138)
        if (o.BySeq && !pattern_is_legal(pt)) { err = "illegal DNA/RNA/Protein sequence: " + pt; return false; }
        std::string q = pt;
        if (o.IgnoreCase)
          for (auto &c : q)
            if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
        kept.push_back(q);
      }
      o.Patterns = kept;
      o.PatternNames = kept;
      break;
    }
    case OP_SUBSEQ:  // bigseqkit-lib/subseq.go:36-96,161
      if (!o.SubseqGtf.empty() || !o.SubseqBed.empty()) {
        code = BSK_ERR_UNSUPPORTED;
        err = "subseq --gtf/--bed are outside the accelerated region path";
        return false;
      }
      if (o.Region.empty()) { err = "one of the options needed: -r/--region, --bed, --gtf"; return false; }
      o.has_region = true;
      if (!parse_region(o.Region, "subseq", o.region_start, o.region_end, err)) return false;
      break;
    case OP_FQ2FA:  // bigseqkit/fq2fa.go:11-18: KitConfig only
      break;
    case OP_DUPLICATE:
      if (o.Times < 0) { err = "times must be >= 0"; return false; }  // make([]string, times) panics below zero
      break;
    case OP_RANGE:
      break;
    default:
      err = "unknown operator";
      return false;
  }
  code = BSK_OK;
  return true;
}

}  // namespace bsk
