// opts.h -- host-side mirror of the reference option structs and alphabets.
//   KitConfig            bigseqkit/helper.go:28-38 (defaults :86-103)
//   SeqOptions           bigseqkit/seq.go:9-55
//   StatsOptions         bigseqkit/stats.go:18-38
//   RmDupOptions         bigseqkit/rmdup.go:13-33
//   TranslateOptions     bigseqkit/translate.go:9-35
//   LocateOptions        bigseqkit/locate.go:9-45
//   GrepOptions          bigseqkit/grep.go:13-49
//   SubseqOptions        bigseqkit/subseq.go:9-35
//   DuplicateOptions     bigseqkit/duplicate.go (Times) ; RangePrepare vars start / end (bigseqkit-lib/range.go:20-24) ;
//   HeadOptions          bigseqkit/head.go:12-22 (N)
// The JSON field names are the Go field names (encoding/json of the struct).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bsk {

enum Op { OP_SEQ, OP_STATS, OP_RMDUP, OP_RMDUP_PREPARE, OP_TRANSLATE, OP_LOCATE, OP_GREP, OP_SUBSEQ, OP_FQ2FA, OP_DUPLICATE, OP_RANGE, OP_INVALID };

// alphabets of shenwei356/bio v0.7.0 seq/alphabet.go as used by the reference
enum Alphabet { AB_NIL = 0, AB_DNA, AB_DNARED, AB_RNA, AB_RNARED, AB_PROTEIN, AB_UNLIMIT, AB_COUNT };
const char *alphabet_name(int a);
// valid[c] != 0 when byte c is a letter, gap or ambiguous symbol of alphabet a
const uint8_t *alphabet_valid(int a);
// complement pair table (identity outside the alphabet's letters)
const uint8_t *alphabet_pair(int a);
// class-mask table for the device alphabet guess: bit k set when byte is valid in {DNA,RNA,DNAred,RNAred,Protein}[k]
void alphabet_class_masks(uint8_t out[256]);
// seq.GuessAlphabetLessConservatively from the AND of class masks over the guessed prefix (empty -> Unlimit)
int alphabet_from_mask(unsigned and_mask, bool empty);
bool pattern_is_legal(const std::string &s);  // valid under DNAredundant, RNAredundant or Protein

struct Opts {
  // KitConfig
  std::string SeqType = "auto";
  int LineWidth = 60;
  std::string IDRegexp = "^(\\S+)\\s?";
  bool IDNCBI = false;
  int AlphabetGuessSeqLength = 10000;
  // SeqOptions
  bool Reverse = false, Complement = false, Name = false, Seq = false, Qual = false, OnlyId = false, RemoveGaps = false;
  std::string GapLetters;  // "- \t." (seq) / "- ." (stats)
  bool LowerCase = false, UpperCase = false, Dna2rna = false, Rna2dna = false, ValidateSeq = false;
  int ValidateSeqLength = 10000, MaxLen = -1, MinLen = -1, QualAsciiBase = 33;
  double MinQual = -1, MaxQual = -1;
  // StatsOptions
  bool Tabular = false, All = false;
  std::string FqEncoding = "sanger";
  // RmDupOptions / GrepOptions
  bool ByName = false, BySeq = false, IgnoreCase = false, OnlyPositiveStrand = false;
  std::string DupSeqsFile, DupNumFile;
  // TranslateOptions
  int TranslTable = 1;
  std::vector<std::string> Frame{"1"};
  bool Trim = false, Clean = false, AllowUnknownCodon = false, InitCodonAsM = false, AppendFrame = false;
  int ListTranslTable = -1, ListTranslTableWithAmbCodons = -1;
  // LocateOptions / GrepOptions
  std::vector<std::string> PatternNames, Patterns;  // locate: name column + sequence; grep: Patterns only
  std::string PatternFile;
  bool Degenerate = false, UseRegexp = false, UseFmi = false, NonGreedy = false, Gtf = false, Bed = false;
  bool HideMatched = false, Circular = false, InvertMatch = false, Count = false, DeleteMatched = false;
  int MaxMismatch = 0;
  // SubseqOptions / Grep region
  std::string Region;
  std::string SubseqGtf, SubseqBed;
  // Duplicate (times), RangePrepare (start, end: 0-based, half-open, as the operator receives them), Head (N)
  int64_t Times = 1, RangeStart = 0, RangeEnd = INT64_MAX, IndexBase = 0;

  // derived in validate()
  int alphabet = AB_NIL;       // from SeqType (AB_NIL = auto)
  int fq_offset = 33;          // stats
  std::vector<int> frames;     // translate
  int region_start = 0, region_end = 0;
  bool has_region = false;
};

Op op_from_name(const char *name);
// Parses the reference JSON for `op` into o and applies the Before() validation of the
// reference operator; returns false with the reference's error text in err.
bool parse_and_validate(Op op, const char *json, Opts &o, std::string &err, int &err_code);

}  // namespace bsk
