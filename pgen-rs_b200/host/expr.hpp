// expr.hpp — the subset of the evalexpr 11.3.0 expression language that pgen-rs exposes
// through --include / --include-var / --include-sam / --fstring.
//
// Reference call sites: /root/reference/src/pfile.rs:87-98 (query_metadata) and :322-328
// (filter_metadata): every column of the row is bound as a String variable, the include
// expression must evaluate to a Boolean and the fstring to a String; any error is an
// `.unwrap()` panic there and PGB_E_EXPR here.
//
// Supported: string literals ("..." with \" and \\ escapes), integer / float / boolean
// literals, identifiers, unary ! and -, binary ^ * / % + - < > <= >= == != && ||,
// parentheses; evalexpr's precedences and eager (non-short-circuit) evaluation.
// Not supported (reported as errors): assignment, ',' tuples, ';' chains, function calls.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <string_view>
#include <vector>

namespace pgb {

struct ExprError {
    std::string msg;
};

struct Value {
    enum Kind { STR, INT, FLOAT, BOOL } kind = BOOL;
    std::string s;
    int64_t i = 0;
    double f = 0;
    bool b = false;
};

class Expr {
  public:
    // Parses `src`; identifiers are resolved against `columns` (name -> field index).
    // Throws ExprError on a syntax error.
    Expr(const std::string &src, const std::vector<std::string> &columns);
    ~Expr();
    Expr(const Expr &) = delete;
    Expr &operator=(const Expr &) = delete;
    // Evaluates on one row (fields indexed like `columns`).  Throws ExprError.
    Value eval(const std::vector<std::string_view> &row) const;
    bool eval_boolean(const std::vector<std::string_view> &row) const;
    std::string eval_string(const std::vector<std::string_view> &row) const;

    struct Node;

  private:
    std::unique_ptr<Node> root_;
};

} // namespace pgb
