// expr.cpp — see expr.hpp.  Tokeniser, precedence-climbing parser and evaluator.
#include "expr.hpp"

#include <math.h>
#include <stdlib.h>

#include <limits>

namespace pgb {

namespace {

enum class Tok { STR, INT, FLOAT, BOOL, IDENT, OP, LPAREN, RPAREN };

struct Token {
    Tok t;
    std::string text; // STR payload, IDENT name, OP spelling
    int64_t i = 0;
    double f = 0;
    bool b = false;
};

bool is_special(char c) {
    switch (c) {
    case '+': case '-': case '*': case '/': case '%': case '^': case '(': case ')': case '=': case '!':
    case '>': case '<': case '&': case '|': case ',': case ';': case '"':
        return true;
    default:
        return false;
    }
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

bool parse_i64(const std::string &w, int64_t *out) {
    if (w.empty()) return false;
    size_t k = 0;
    if (w.size() > 2 + k && w[k] == '0' && (w[k + 1] == 'x')) {
        // evalexpr's parse_dec_or_hex: 0x-prefixed hexadecimal
        uint64_t v = 0;
        for (size_t j = k + 2; j < w.size(); j++) {
            char c = w[j];
            int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
            if (d < 0) return false;
            if (v >> 59) return false;
            v = v * 16 + (uint64_t)d;
        }
        if (v > (uint64_t)std::numeric_limits<int64_t>::max()) return false;
        *out = (int64_t)v;
        return true;
    }
    if (k == w.size()) return false;
    uint64_t v = 0;
    for (size_t j = k; j < w.size(); j++) {
        if (w[j] < '0' || w[j] > '9') return false;
        if (v > (std::numeric_limits<uint64_t>::max() - 9) / 10) return false;
        v = v * 10 + (uint64_t)(w[j] - '0');
    }
    if (v > (uint64_t)std::numeric_limits<int64_t>::max()) return false;
    *out = (int64_t)v;
    return true;
}

bool parse_f64(const std::string &w, double *out) {
    // Decimal floats only ("1.5", ".5", "1e-3"); Rust's additional spellings (inf, nan)
    // are deliberately treated as identifiers here.
    if (w.empty()) return false;
    bool digit = false;
    for (char c : w) {
        if (c >= '0' && c <= '9') digit = true;
        else if (c != '.' && c != 'e' && c != 'E' && c != '+' && c != '-') return false;
    }
    if (!digit) return false;
    char *e = nullptr;
    double v = strtod(w.c_str(), &e);
    if (!e || *e) return false;
    *out = v;
    return true;
}

std::vector<Token> tokenize(const std::string &src) {
    std::vector<Token> out;
    size_t i = 0, n = src.size();
    while (i < n) {
        char c = src[i];
        if (is_space(c)) { i++; continue; }
        if (c == '"') {
            i++;
            Token t{Tok::STR, {}};
            for (;;) {
                if (i >= n) throw ExprError{"unmatched double quote"};
                c = src[i];
                if (c == '\\') {
                    if (i + 1 < n && (src[i + 1] == '"' || src[i + 1] == '\\')) { t.text.push_back(src[i + 1]); i += 2; }
                    else throw ExprError{"illegal escape sequence"};
                } else if (c == '"') { i++; break; }
                else { t.text.push_back(c); i++; }
            }
            out.push_back(std::move(t));
            continue;
        }
        if (c == '(') { out.push_back(Token{Tok::LPAREN, "("}); i++; continue; }
        if (c == ')') { out.push_back(Token{Tok::RPAREN, ")"}); i++; continue; }
        if (is_special(c)) {
            char d = i + 1 < n ? src[i + 1] : 0;
            std::string op(1, c);
            if ((c == '=' || c == '!' || c == '<' || c == '>') && d == '=') op.push_back('=');
            else if ((c == '&' && d == '&') || (c == '|' && d == '|')) op.push_back(d);
            if (op == "=" || op == "&" || op == "|" || op == "," || op == ";")
                throw ExprError{"unsupported operator '" + op + "'"};
            i += op.size();
            out.push_back(Token{Tok::OP, op});
            continue;
        }
        size_t j = i;
        while (j < n && !is_special(src[j]) && !is_space(src[j])) j++;
        std::string w = src.substr(i, j - i);
        i = j;
        Token t{Tok::IDENT, w};
        if (parse_i64(w, &t.i)) t.t = Tok::INT;
        else if (parse_f64(w, &t.f)) t.t = Tok::FLOAT;
        else if (w == "true") { t.t = Tok::BOOL; t.b = true; }
        else if (w == "false") { t.t = Tok::BOOL; t.b = false; }
        out.push_back(std::move(t));
    }
    return out;
}

int bin_prec(const std::string &op) {
    if (op == "^") return 120;
    if (op == "*" || op == "/" || op == "%") return 100;
    if (op == "+" || op == "-") return 95;
    if (op == "<" || op == ">" || op == "<=" || op == ">=" || op == "==" || op == "!=") return 80;
    if (op == "&&") return 75;
    if (op == "||") return 70;
    return -1;
}

} // namespace

struct Expr::Node {
    enum Kind { LIT, VAR, UNARY, BINARY } kind;
    Value lit;
    int col = -1;
    std::string name;
    std::string op;
    std::unique_ptr<Node> a, b;
};

namespace {

struct Parser {
    const std::vector<Token> &t;
    const std::vector<std::string> &cols;
    size_t i = 0;

    const Token *peek() const { return i < t.size() ? &t[i] : nullptr; }

    std::unique_ptr<Expr::Node> expr(int min_prec) {
        auto lhs = unary();
        for (;;) {
            const Token *tk = peek();
            if (!tk || tk->t != Tok::OP) return lhs;
            int prec = bin_prec(tk->text);
            if (prec < 0 || prec < min_prec) return lhs;
            std::string op = tk->text;
            i++;
            auto rhs = expr(op == "^" ? prec : prec + 1);
            auto n = std::make_unique<Expr::Node>();
            n->kind = Expr::Node::BINARY;
            n->op = op;
            n->a = std::move(lhs);
            n->b = std::move(rhs);
            lhs = std::move(n);
        }
    }

    std::unique_ptr<Expr::Node> unary() {
        const Token *tk = peek();
        if (!tk) throw ExprError{"unexpected end of expression"};
        if (tk->t == Tok::OP && (tk->text == "!" || tk->text == "-")) {
            std::string op = tk->text;
            i++;
            auto n = std::make_unique<Expr::Node>();
            n->kind = Expr::Node::UNARY;
            n->op = op;
            n->a = expr(110);
            return n;
        }
        if (tk->t == Tok::LPAREN) {
            i++;
            auto n = expr(0);
            const Token *nx = peek();
            if (!nx || nx->t != Tok::RPAREN) throw ExprError{"unmatched parenthesis"};
            i++;
            return n;
        }
        auto n = std::make_unique<Expr::Node>();
        switch (tk->t) {
        case Tok::STR: n->kind = Expr::Node::LIT; n->lit.kind = Value::STR; n->lit.s = tk->text; break;
        case Tok::INT: n->kind = Expr::Node::LIT; n->lit.kind = Value::INT; n->lit.i = tk->i; break;
        case Tok::FLOAT: n->kind = Expr::Node::LIT; n->lit.kind = Value::FLOAT; n->lit.f = tk->f; break;
        case Tok::BOOL: n->kind = Expr::Node::LIT; n->lit.kind = Value::BOOL; n->lit.b = tk->b; break;
        case Tok::IDENT: {
            n->kind = Expr::Node::VAR;
            n->name = tk->text;
            // HashMapContext::set_value overwrites: with duplicate column names the last wins
            for (size_t c = 0; c < cols.size(); c++)
                if (cols[c] == tk->text) n->col = (int)c;
            break;
        }
        default: throw ExprError{"unexpected token '" + tk->text + "'"};
        }
        i++;
        const Token *nx = peek();
        if (n->kind == Expr::Node::VAR && nx && (nx->t == Tok::LPAREN || nx->t == Tok::STR || nx->t == Tok::INT ||
                                                  nx->t == Tok::FLOAT || nx->t == Tok::BOOL || nx->t == Tok::IDENT))
            throw ExprError{"function calls are not supported ('" + n->name + "')"};
        return n;
    }
};

bool is_num(const Value &v) { return v.kind == Value::INT || v.kind == Value::FLOAT; }
double as_f(const Value &v) { return v.kind == Value::INT ? (double)v.i : v.f; }

Value mk_bool(bool b) { Value v; v.kind = Value::BOOL; v.b = b; return v; }
Value mk_int(int64_t i) { Value v; v.kind = Value::INT; v.i = i; return v; }
Value mk_float(double f) { Value v; v.kind = Value::FLOAT; v.f = f; return v; }

bool value_eq(const Value &a, const Value &b) {
    if (a.kind != b.kind) return false; // evalexpr derives PartialEq on Value: Int(1) != Float(1.0)
    switch (a.kind) {
    case Value::STR: return a.s == b.s;
    case Value::INT: return a.i == b.i;
    case Value::FLOAT: return a.f == b.f;
    default: return a.b == b.b;
    }
}

Value eval_node(const Expr::Node *n, const std::vector<std::string_view> &row) {
    switch (n->kind) {
    case Expr::Node::LIT: return n->lit;
    case Expr::Node::VAR: {
        if (n->col < 0 || (size_t)n->col >= row.size()) throw ExprError{"variable identifier '" + n->name + "' not found"};
        Value v;
        v.kind = Value::STR;
        v.s.assign(row[n->col].data(), row[n->col].size());
        return v;
    }
    case Expr::Node::UNARY: {
        Value a = eval_node(n->a.get(), row);
        if (n->op == "!") {
            if (a.kind != Value::BOOL) throw ExprError{"expected a boolean"};
            return mk_bool(!a.b);
        }
        if (a.kind == Value::INT) {
            if (a.i == std::numeric_limits<int64_t>::min()) throw ExprError{"negation overflow"};
            return mk_int(-a.i);
        }
        if (a.kind == Value::FLOAT) return mk_float(-a.f);
        throw ExprError{"expected a number"};
    }
    default: break;
    }
    // evalexpr evaluates every argument before applying the operator (no short circuit)
    Value a = eval_node(n->a.get(), row);
    Value b = eval_node(n->b.get(), row);
    const std::string &op = n->op;
    if (op == "&&" || op == "||") {
        if (a.kind != Value::BOOL || b.kind != Value::BOOL) throw ExprError{"expected a boolean"};
        return mk_bool(op == "&&" ? (a.b && b.b) : (a.b || b.b));
    }
    if (op == "==") return mk_bool(value_eq(a, b));
    if (op == "!=") return mk_bool(!value_eq(a, b));
    if (op == "<" || op == ">" || op == "<=" || op == ">=") {
        int cmp;
        if (a.kind == Value::STR && b.kind == Value::STR) cmp = a.s < b.s ? -1 : (a.s == b.s ? 0 : 1);
        else if (a.kind == Value::INT && b.kind == Value::INT) cmp = a.i < b.i ? -1 : (a.i == b.i ? 0 : 1);
        else if (is_num(a) && is_num(b)) {
            double x = as_f(a), y = as_f(b);
            if (x != x || y != y) return mk_bool(false);
            cmp = x < y ? -1 : (x == y ? 0 : 1);
        } else throw ExprError{"expected two numbers or two strings"};
        if (op == "<") return mk_bool(cmp < 0);
        if (op == ">") return mk_bool(cmp > 0);
        if (op == "<=") return mk_bool(cmp <= 0);
        return mk_bool(cmp >= 0);
    }
    if (op == "+" && a.kind == Value::STR && b.kind == Value::STR) {
        Value v;
        v.kind = Value::STR;
        v.s = a.s + b.s;
        return v;
    }
    if (!is_num(a) || !is_num(b)) throw ExprError{op == "+" ? "expected two numbers or two strings" : "expected a number"};
    if (op == "^") return mk_float(pow(as_f(a), as_f(b)));
    if (a.kind == Value::INT && b.kind == Value::INT) {
        int64_t r;
        if (op == "+") { if (__builtin_add_overflow(a.i, b.i, &r)) throw ExprError{"addition overflow"}; return mk_int(r); }
        if (op == "-") { if (__builtin_sub_overflow(a.i, b.i, &r)) throw ExprError{"subtraction overflow"}; return mk_int(r); }
        if (op == "*") { if (__builtin_mul_overflow(a.i, b.i, &r)) throw ExprError{"multiplication overflow"}; return mk_int(r); }
        if (b.i == 0 || (a.i == std::numeric_limits<int64_t>::min() && b.i == -1)) throw ExprError{"division error"};
        return mk_int(op == "/" ? a.i / b.i : a.i % b.i);
    }
    double x = as_f(a), y = as_f(b);
    if (op == "+") return mk_float(x + y);
    if (op == "-") return mk_float(x - y);
    if (op == "*") return mk_float(x * y);
    if (op == "/") return mk_float(x / y);
    return mk_float(fmod(x, y));
}

} // namespace

Expr::Expr(const std::string &src, const std::vector<std::string> &columns) {
    std::vector<Token> toks = tokenize(src);
    if (toks.empty()) throw ExprError{"empty expression"};
    Parser p{toks, columns};
    root_ = p.expr(0);
    if (p.i != toks.size()) throw ExprError{"unexpected token '" + toks[p.i].text + "'"};
}

Expr::~Expr() = default;

Value Expr::eval(const std::vector<std::string_view> &row) const { return eval_node(root_.get(), row); }

bool Expr::eval_boolean(const std::vector<std::string_view> &row) const {
    Value v = eval(row);
    if (v.kind != Value::BOOL) throw ExprError{"expected a boolean"};
    return v.b;
}

std::string Expr::eval_string(const std::vector<std::string_view> &row) const {
    Value v = eval(row);
    if (v.kind != Value::STR) throw ExprError{"expected a string"};
    return v.s;
}

} // namespace pgb
