// cli.cpp — `pgen-b200`: flag-for-flag mirror of the reference CLI
// (/root/reference/src/cli.rs:5-62, /root/reference/src/main.rs:92-127):
//
//   pgen-b200 query  <PFILE_PREFIX> -f/--fstring <EXPR> [-i/--include <EXPR>] [-s/--samples]
//   pgen-b200 filter <PFILE_PREFIX> [--include-var <EXPR>] [--include-sam <EXPR>] [-o/--out <PATH>]
//
// plus one extension: --devices 0,1,.. (filter) to shard variant ranges over several GPUs.
// Failures exit with status 101, the exit code of the reference's panics.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <optional>
#include <string>
#include <vector>

#include "pfile.hpp"

namespace {

const char *kVersion = "0.1.0";

void usage(FILE *o) {
    fputs("Usage: pgen-b200 <COMMAND>\n\n"
          "Commands:\n"
          "  query   Queries the pgen, outputting to stdout\n"
          "  filter  Filters the pgen, outputting to a VCF\n"
          "  help    Print this message or the help of the given subcommand(s)\n\n"
          "Options:\n"
          "  -h, --help     Print help\n"
          "  -V, --version  Print version\n",
          o);
}

void usage_query(FILE *o) {
    fputs("Usage: pgen-b200 query [OPTIONS] --fstring <QUERY_FSTRING> <PFILE_PREFIX>\n\n"
          "Arguments:\n  <PFILE_PREFIX>  The prefix of the pgen file triples\n\n"
          "Options:\n"
          "  -f, --fstring <QUERY_FSTRING>  An expression specifying what to output to stdout\n"
          "  -i, --include <QUERY>          An expression specifying which variants (default) or samples (if -s is passed) to keep\n"
          "  -s, --samples                  When passed, the query is over the samples\n"
          "  -h, --help                     Print help\n",
          o);
}

void usage_filter(FILE *o) {
    fputs("Usage: pgen-b200 filter [OPTIONS] <PFILE_PREFIX>\n\n"
          "Arguments:\n  <PFILE_PREFIX>  The prefix of the pgen file triples\n\n"
          "Options:\n"
          "      --include-var <VAR_QUERY>  An expression specifying which variants to keep. If not passed, keeps all variants\n"
          "      --include-sam <SAM_QUERY>  An expression specifying which samples to keep. If not passed, keeps all samples\n"
          "  -o, --out <OUT_FILE>           The output file name (defaults to PFILE_PREFIX.pgen-rs.vcf)\n"
          "      --devices <IDS>            Comma-separated CUDA device ids to shard variant ranges over (default 0)\n"
          "  -h, --help                     Print help\n",
          o);
}

// clap accepts `--flag value`, `--flag=value`, `-f value`, `-fvalue`.
bool take(const std::vector<std::string> &a, size_t &i, const char *lng, const char *sht, std::string *out) {
    const std::string &s = a[i];
    if (lng) {
        std::string l = std::string("--") + lng;
        if (s == l) {
            if (i + 1 >= a.size()) { fprintf(stderr, "error: a value is required for '%s <..>' but none was supplied\n", l.c_str()); exit(2); }
            *out = a[++i];
            return true;
        }
        if (s.rfind(l + "=", 0) == 0) { *out = s.substr(l.size() + 1); return true; }
    }
    if (sht) {
        std::string h = std::string("-") + sht;
        if (s == h) {
            if (i + 1 >= a.size()) { fprintf(stderr, "error: a value is required for '%s <..>' but none was supplied\n", h.c_str()); exit(2); }
            *out = a[++i];
            return true;
        }
        if (s.size() > 2 && s.rfind(h, 0) == 0 && s[1] != '-') { *out = s.substr(s[2] == '=' ? 3 : 2); return true; }
    }
    return false;
}

} // namespace

int main(int argc, char **argv) {
    std::vector<std::string> a(argv + 1, argv + argc);
    if (a.empty()) { usage(stderr); return 2; }
    const std::string cmd = a[0];
    if (cmd == "-h" || cmd == "--help" || cmd == "help") { usage(stdout); return 0; }
    if (cmd == "-V" || cmd == "--version") { printf("pgen-b200 %s\n", kVersion); return 0; }
    try {
        if (cmd == "query") {
            std::optional<std::string> prefix, fstring, query;
            bool samples = false;
            for (size_t i = 1; i < a.size(); i++) {
                std::string v;
                if (a[i] == "-h" || a[i] == "--help") { usage_query(stdout); return 0; }
                if (take(a, i, "fstring", "f", &v)) fstring = v;
                else if (take(a, i, "include", "i", &v)) query = v;
                else if (a[i] == "-s" || a[i] == "--samples") samples = true;
                else if (a[i].size() > 1 && a[i][0] == '-') { fprintf(stderr, "error: unexpected argument '%s' found\n", a[i].c_str()); return 2; }
                else if (!prefix) prefix = a[i];
                else { fprintf(stderr, "error: unexpected argument '%s' found\n", a[i].c_str()); return 2; }
            }
            if (!prefix || !fstring) { fputs("error: the following required arguments were not provided\n", stderr); usage_query(stderr); return 2; }
            pgb::Pfile pf = pgb::Pfile::from_prefix(*prefix); // main.rs:101
            pgb::MetaTable t = samples ? pf.psam_reader() : pf.pvar_reader();
            pf.query_metadata(t, query, *fstring, 1);
            return 0;
        }
        if (cmd == "filter") {
            std::optional<std::string> prefix, var_query, sam_query, out;
            std::vector<int> devices;
            for (size_t i = 1; i < a.size(); i++) {
                std::string v;
                if (a[i] == "-h" || a[i] == "--help") { usage_filter(stdout); return 0; }
                if (take(a, i, "include-var", nullptr, &v)) var_query = v;
                else if (take(a, i, "include-sam", nullptr, &v)) sam_query = v;
                else if (take(a, i, "out", "o", &v)) out = v;
                else if (take(a, i, "devices", nullptr, &v)) {
                    char *s = v.data();
                    for (char *tok = strtok(s, ","); tok; tok = strtok(nullptr, ",")) devices.push_back(atoi(tok));
                } else if (a[i].size() > 1 && a[i][0] == '-') { fprintf(stderr, "error: unexpected argument '%s' found\n", a[i].c_str()); return 2; }
                else if (!prefix) prefix = a[i];
                else { fprintf(stderr, "error: unexpected argument '%s' found\n", a[i].c_str()); return 2; }
            }
            if (!prefix) { fputs("error: the following required arguments were not provided:\n  <PFILE_PREFIX>\n", stderr); usage_filter(stderr); return 2; }
            pgb::Pfile pf = pgb::Pfile::from_prefix(*prefix);                       // main.rs:120
            std::string out_file = out ? *out : pf.pfile_prefix + ".pgen-rs.vcf"; // main.rs:121-122
            pf.output_vcf(sam_query, var_query, out_file, devices.empty() ? nullptr : devices.data(), (int)devices.size());
            return 0;
        }
    } catch (const pgb::PfileError &e) {
        fprintf(stderr, "pgen-b200: %s (%s)\n", e.msg.c_str(), pgb_strerror(e.status));
        return 101; // a Rust panic exits with 101
    }
    fprintf(stderr, "error: unrecognized subcommand '%s'\n", cmd.c_str());
    usage(stderr);
    return 2;
}
