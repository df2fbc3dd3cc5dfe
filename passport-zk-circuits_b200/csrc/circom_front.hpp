// circom_front.hpp - lexer + recursive-descent parser for the circom 2.1.x subset used by
// passport-zk-circuits.  Produces a plain pointer AST.
// Product code: shares nothing with the CPU oracle under oracle/.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "u256.hpp"

namespace pzk {

struct CompileError : std::runtime_error {
  explicit CompileError(const std::string& m) : std::runtime_error(m) {}
};

// ---- name interning --------------------------------------------------------
struct Interner {
  std::unordered_map<std::string, int> ids;
  std::vector<std::string> names;
  int get(const std::string& s) {
    auto it = ids.find(s);
    if (it != ids.end()) return it->second;
    int id = (int)names.size();
    ids.emplace(s, id);
    names.push_back(s);
    return id;
  }
  const std::string& str(int id) const { return names[id]; }
};

// ---- tokens ---------------------------------------------------------------------
enum Tok { T_EOF, T_NUM, T_ID, T_STR, T_OP };

enum OpCode {
  O_NONE = 0,
  O_ADD, O_SUB, O_MUL, O_DIV, O_IDIV, O_MOD, O_POW,
  O_SHL, O_SHR, O_BAND, O_BOR, O_BXOR,
  O_EQ, O_NE, O_LT, O_GT, O_LE, O_GE, O_LAND, O_LOR,
  O_NEG, O_LNOT, O_BNOT,
  // punctuation / assignment (lexer only)
  O_ASSIGN, O_CONSTRAIN_L, O_CONSTRAIN_R, O_WITNESS_L, O_WITNESS_R, O_CEQ,
  O_ADD_A, O_SUB_A, O_MUL_A, O_DIV_A, O_IDIV_A, O_MOD_A, O_SHL_A, O_SHR_A, O_BAND_A, O_BOR_A,
  O_BXOR_A, O_POW_A, O_INC, O_DEC,
  O_QUEST, O_COLON, O_SEMI, O_COMMA, O_DOT, O_LP, O_RP, O_LB, O_RB, O_LC, O_RC
};

struct Token {
  Tok t;
  int op;        // OpCode for T_OP
  int id;        // interned name for T_ID / string for T_STR (index into strs)
  U256 num;
  int line;
};

struct Expr {
  enum K { NUM, VAR, IDX, MEM, CALL, ARR, UN, BIN, TERN } k;
  int op = 0;
  int name = -1;
  U256 num;
  Expr *a = nullptr, *b = nullptr, *c = nullptr;
  std::vector<Expr*> args;
  int line = 0;
};

struct DeclItem {
  int name;
  std::vector<Expr*> dims;
  int init_op = 0;  // O_CONSTRAIN_L / O_WITNESS_L / O_ASSIGN
  Expr* init = nullptr;
};

struct Stmt {
  enum K { BLOCK, SIGDECL, VARDECL, COMPDECL, ASSIGN, CONSTR, IF, FOR, WHILE, RETURN, ASSERT, LOG, INCDEC } k;
  int op = 0;       // ASSIGN: assignment operator; SIGDECL: 0 mid, 1 in, 2 out; INCDEC: +1/-1
  Expr *lhs = nullptr, *rhs = nullptr, *cond = nullptr;
  Stmt *s1 = nullptr, *s2 = nullptr, *s3 = nullptr, *body = nullptr;
  std::vector<Stmt*> stmts;
  std::vector<DeclItem> items;
  int line = 0;
  int file = 0;
  // analysis cache
  int has_decl = -1;  // contains declarations / component instantiation (phase A relevant)
};

struct TemplateDef {
  int name;
  std::vector<int> params;
  Stmt* body;
  int file, line;
};

struct MainDef {
  bool present = false;
  std::vector<int> publics;
  Expr* call = nullptr;
};

struct SourceUnit {
  Interner names;
  std::vector<std::string> files;
  std::unordered_map<int, TemplateDef> templates;
  std::unordered_map<int, TemplateDef> functions;
  MainDef main;
  std::vector<std::string> missing;
  std::set<std::string> seen;
  std::vector<std::unique_ptr<Expr>> expr_pool;
  std::vector<std::unique_ptr<Stmt>> stmt_pool;
  Expr* new_expr(Expr::K k, int line) {
    expr_pool.emplace_back(new Expr());
    Expr* e = expr_pool.back().get();
    e->k = k; e->line = line;
    return e;
  }
  Stmt* new_stmt(Stmt::K k, int line, int file) {
    stmt_pool.emplace_back(new Stmt());
    Stmt* s = stmt_pool.back().get();
    s->k = k; s->line = line; s->file = file;
    return s;
  }
};

// ---- lexer -----------------------------------------------------------------------
class Lexer {
 public:
  Lexer(const std::string& src, const std::string& fname, SourceUnit& u) : s(src), f(fname), unit(u) {}
  std::vector<Token> run(std::vector<std::string>& strs) {
    std::vector<Token> out;
    size_t n = s.size(), i = 0;
    int line = 1;
    auto push_op = [&](int op, int len) {
      Token t; t.t = T_OP; t.op = op; t.id = -1; t.line = line; out.push_back(t); i += len;
    };
    while (i < n) {
      char c = s[i];
      if (c == '\n') { line++; i++; continue; }
      if (c == ' ' || c == '\t' || c == '\r') { i++; continue; }
      if (c == '/' && i + 1 < n && s[i + 1] == '/') { while (i < n && s[i] != '\n') i++; continue; }
      if (c == '/' && i + 1 < n && s[i + 1] == '*') {
        i += 2;
        while (i + 1 < n && !(s[i] == '*' && s[i + 1] == '/')) { if (s[i] == '\n') line++; i++; }
        i += 2; continue;
      }
      if (c >= '0' && c <= '9') {
        size_t j = i;
        if (c == '0' && i + 1 < n && (s[i + 1] == 'x' || s[i + 1] == 'X')) {
          j = i + 2;
          while (j < n && isxdigit((unsigned char)s[j])) j++;
        } else {
          while (j < n && s[j] >= '0' && s[j] <= '9') j++;
        }
        Token t; t.t = T_NUM; t.op = 0; t.id = -1; t.line = line;
        t.num = parse_number(s.data() + i, j - i);
        out.push_back(t); i = j; continue;
      }
      if (isalpha((unsigned char)c) || c == '_' || c == '$') {
        size_t j = i + 1;
        while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_' || s[j] == '$')) j++;
        Token t; t.t = T_ID; t.op = 0; t.line = line;
        t.id = unit.names.get(s.substr(i, j - i));
        out.push_back(t); i = j; continue;
      }
      if (c == '"') {
        size_t j = i + 1;
        while (j < n && s[j] != '"') { if (s[j] == '\\') j++; j++; }
        Token t; t.t = T_STR; t.op = 0; t.line = line; t.id = (int)strs.size();
        strs.push_back(s.substr(i + 1, j - i - 1));
        out.push_back(t); i = j + 1; continue;
      }
      auto at = [&](const char* lit) { size_t L = strlen(lit); return s.compare(i, L, lit) == 0; };
      // three-char operators first
      if (at("<==")) { push_op(O_CONSTRAIN_L, 3); continue; }
      if (at("==>")) { push_op(O_CONSTRAIN_R, 3); continue; }
      if (at("<--")) { push_op(O_WITNESS_L, 3); continue; }
      if (at("-->")) { push_op(O_WITNESS_R, 3); continue; }
      if (at("===")) { push_op(O_CEQ, 3); continue; }
      if (at("**=")) { push_op(O_POW_A, 3); continue; }
      if (at("<<=")) { push_op(O_SHL_A, 3); continue; }
      if (at(">>=")) { push_op(O_SHR_A, 3); continue; }
      if (at("**")) { push_op(O_POW, 2); continue; }
      if (at("++")) { push_op(O_INC, 2); continue; }
      if (at("--")) { push_op(O_DEC, 2); continue; }
      if (at("&&")) { push_op(O_LAND, 2); continue; }
      if (at("||")) { push_op(O_LOR, 2); continue; }
      if (at("==")) { push_op(O_EQ, 2); continue; }
      if (at("!=")) { push_op(O_NE, 2); continue; }
      if (at("<=")) { push_op(O_LE, 2); continue; }
      if (at(">=")) { push_op(O_GE, 2); continue; }
      if (at("<<")) { push_op(O_SHL, 2); continue; }
      if (at(">>")) { push_op(O_SHR, 2); continue; }
      if (at("+=")) { push_op(O_ADD_A, 2); continue; }
      if (at("-=")) { push_op(O_SUB_A, 2); continue; }
      if (at("*=")) { push_op(O_MUL_A, 2); continue; }
      if (at("/=")) { push_op(O_DIV_A, 2); continue; }
      if (at("\\=")) { push_op(O_IDIV_A, 2); continue; }
      if (at("%=")) { push_op(O_MOD_A, 2); continue; }
      if (at("&=")) { push_op(O_BAND_A, 2); continue; }
      if (at("|=")) { push_op(O_BOR_A, 2); continue; }
      if (at("^=")) { push_op(O_BXOR_A, 2); continue; }
      switch (c) {
        case '+': push_op(O_ADD, 1); continue;
        case '-': push_op(O_SUB, 1); continue;
        case '*': push_op(O_MUL, 1); continue;
        case '/': push_op(O_DIV, 1); continue;
        case '\\': push_op(O_IDIV, 1); continue;
        case '%': push_op(O_MOD, 1); continue;
        case '<': push_op(O_LT, 1); continue;
        case '>': push_op(O_GT, 1); continue;
        case '=': push_op(O_ASSIGN, 1); continue;
        case '!': push_op(O_LNOT, 1); continue;
        case '~': push_op(O_BNOT, 1); continue;
        case '&': push_op(O_BAND, 1); continue;
        case '|': push_op(O_BOR, 1); continue;
        case '^': push_op(O_BXOR, 1); continue;
        case '?': push_op(O_QUEST, 1); continue;
        case ':': push_op(O_COLON, 1); continue;
        case ';': push_op(O_SEMI, 1); continue;
        case ',': push_op(O_COMMA, 1); continue;
        case '.': push_op(O_DOT, 1); continue;
        case '(': push_op(O_LP, 1); continue;
        case ')': push_op(O_RP, 1); continue;
        case '[': push_op(O_LB, 1); continue;
        case ']': push_op(O_RB, 1); continue;
        case '{': push_op(O_LC, 1); continue;
        case '}': push_op(O_RC, 1); continue;
      }
      throw CompileError(f + ":" + std::to_string(line) + ": bad character '" + std::string(1, c) + "'");
    }
    Token t; t.t = T_EOF; t.op = 0; t.id = -1; t.line = line; out.push_back(t);
    return out;
  }

 private:
  const std::string& s;
  std::string f;
  SourceUnit& unit;
};

// ---- parser ----------------------------------------------------------------------
class Parser {
 public:
  Parser(SourceUnit& u, const std::string& path, int file_id) : unit(u), fname(path), file(file_id) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw CompileError("cannot open " + path);
    std::stringstream ss; ss << in.rdbuf();
    src = ss.str();
    Lexer lx(src, path, unit);
    toks = lx.run(strs);
    auto& N = unit.names;
    kw_pragma = N.get("pragma"); kw_include = N.get("include"); kw_template = N.get("template");
    kw_function = N.get("function"); kw_component = N.get("component"); kw_signal = N.get("signal");
    kw_var = N.get("var"); kw_input = N.get("input"); kw_output = N.get("output");
    kw_public = N.get("public"); kw_if = N.get("if"); kw_else = N.get("else"); kw_for = N.get("for");
    kw_while = N.get("while"); kw_return = N.get("return"); kw_assert = N.get("assert");
    kw_log = N.get("log"); kw_main = N.get("main"); kw_parallel = N.get("parallel");
    kw_custom = N.get("custom");
  }

  // returns the include paths (as written)
  std::vector<std::string> parse_file(bool is_root) {
    std::vector<std::string> incs;
    while (cur().t != T_EOF) {
      if (is_kw(kw_pragma)) { while (!eat(O_SEMI)) p++; continue; }
      if (is_kw(kw_include)) {
        p++;
        if (cur().t != T_STR) err("expected include path");
        incs.push_back(strs[cur().id]); p++;
        expect(O_SEMI); continue;
      }
      if (is_kw(kw_template) || is_kw(kw_function)) {
        bool is_t = is_kw(kw_template);
        p++;
        while (is_kw(kw_custom) || is_kw(kw_parallel)) p++;
        TemplateDef d; d.line = cur().line; d.file = file;
        d.name = ident();
        expect(O_LP);
        while (!eat(O_RP)) { d.params.push_back(ident()); eat(O_COMMA); }
        d.body = block();
        if (is_t) unit.templates[d.name] = d; else unit.functions[d.name] = d;
        continue;
      }
      if (is_kw(kw_component)) {
        p++;
        if (!is_kw(kw_main)) err("only `component main` is allowed at top level");
        p++;
        MainDef m; m.present = true;
        if (eat(O_LC)) {
          if (!is_kw(kw_public)) err("expected public");
          p++; expect(O_LB);
          while (!eat(O_RB)) { m.publics.push_back(ident()); eat(O_COMMA); }
          expect(O_RC);
        }
        expect(O_ASSIGN);
        m.call = expr();
        expect(O_SEMI);
        if (is_root) unit.main = m;
        continue;
      }
      err("unexpected top-level token");
    }
    return incs;
  }

 private:
  SourceUnit& unit;
  std::string fname, src;
  int file;
  std::vector<Token> toks;
  std::vector<std::string> strs;
  size_t p = 0;
  int kw_pragma, kw_include, kw_template, kw_function, kw_component, kw_signal, kw_var, kw_input,
      kw_output, kw_public, kw_if, kw_else, kw_for, kw_while, kw_return, kw_assert, kw_log, kw_main,
      kw_parallel, kw_custom;

  const Token& cur() const { return toks[p]; }
  const Token& peek(int k) const { return toks[std::min(p + k, toks.size() - 1)]; }
  bool is_kw(int kw) const { return cur().t == T_ID && cur().id == kw; }
  bool is_op(int op) const { return cur().t == T_OP && cur().op == op; }
  bool eat(int op) { if (is_op(op)) { p++; return true; } return false; }
  [[noreturn]] void err(const std::string& m) {
    throw CompileError(fname + ":" + std::to_string(cur().line) + ": " + m);
  }
  void expect(int op) { if (!eat(op)) err("syntax error (expected punctuation " + std::to_string(op) + ")"); }
  int ident() { if (cur().t != T_ID) err("expected identifier"); return toks[p++].id; }

  Stmt* block() {
    Stmt* b = unit.new_stmt(Stmt::BLOCK, cur().line, file);
    expect(O_LC);
    while (!eat(O_RC)) b->stmts.push_back(stmt());
    return b;
  }

  std::vector<Expr*> dims() {
    std::vector<Expr*> d;
    while (eat(O_LB)) { d.push_back(expr()); expect(O_RB); }
    return d;
  }

  Stmt* stmt() {
    if (is_op(O_LC)) return block();
    if (cur().t == T_ID) {
      int kw = cur().id;
      if (kw == kw_signal) { Stmt* s = decl_signal(); expect(O_SEMI); return s; }
      if (kw == kw_var) { Stmt* s = decl_var(); expect(O_SEMI); return s; }
      if (kw == kw_component) { Stmt* s = decl_comp(); expect(O_SEMI); return s; }
      if (kw == kw_if) {
        Stmt* s = unit.new_stmt(Stmt::IF, cur().line, file);
        p++; expect(O_LP); s->cond = expr(); expect(O_RP);
        s->s1 = stmt();
        if (is_kw(kw_else)) { p++; s->s2 = stmt(); }
        return s;
      }
      if (kw == kw_for) {
        Stmt* s = unit.new_stmt(Stmt::FOR, cur().line, file);
        p++; expect(O_LP);
        s->s1 = simple_stmt(); expect(O_SEMI);
        s->cond = expr(); expect(O_SEMI);
        s->s2 = simple_stmt(); expect(O_RP);
        s->body = stmt();
        return s;
      }
      if (kw == kw_while) {
        Stmt* s = unit.new_stmt(Stmt::WHILE, cur().line, file);
        p++; expect(O_LP); s->cond = expr(); expect(O_RP);
        s->body = stmt();
        return s;
      }
      if (kw == kw_return) {
        Stmt* s = unit.new_stmt(Stmt::RETURN, cur().line, file);
        p++; s->rhs = expr(); expect(O_SEMI);
        return s;
      }
      if (kw == kw_assert) {
        Stmt* s = unit.new_stmt(Stmt::ASSERT, cur().line, file);
        p++; expect(O_LP); s->cond = expr(); expect(O_RP); expect(O_SEMI);
        return s;
      }
      if (kw == kw_log) {
        Stmt* s = unit.new_stmt(Stmt::LOG, cur().line, file);
        p++; expect(O_LP);
        int depth = 1;
        while (depth) { if (is_op(O_LP)) depth++; else if (is_op(O_RP)) depth--; if (cur().t == T_EOF) err("eof in log"); p++; }
        expect(O_SEMI);
        return s;
      }
    }
    Stmt* s = simple_stmt();
    expect(O_SEMI);
    return s;
  }

  static int compound_base(int op) {
    switch (op) {
      case O_ADD_A: return O_ADD; case O_SUB_A: return O_SUB; case O_MUL_A: return O_MUL;
      case O_DIV_A: return O_DIV; case O_IDIV_A: return O_IDIV; case O_MOD_A: return O_MOD;
      case O_SHL_A: return O_SHL; case O_SHR_A: return O_SHR; case O_BAND_A: return O_BAND;
      case O_BOR_A: return O_BOR; case O_BXOR_A: return O_BXOR; case O_POW_A: return O_POW;
    }
    return 0;
  }

  Stmt* simple_stmt() {
    if (is_kw(kw_var)) return decl_var();
    int line = cur().line;
    if (is_op(O_INC) || is_op(O_DEC)) {
      int d = is_op(O_INC) ? 1 : -1; p++;
      Stmt* s = unit.new_stmt(Stmt::INCDEC, line, file);
      s->lhs = expr(); s->op = d; return s;
    }
    Expr* e = expr();
    if (cur().t == T_OP) {
      int op = cur().op;
      if (op == O_ASSIGN || op == O_CONSTRAIN_L || op == O_WITNESS_L) {
        p++;
        Stmt* s = unit.new_stmt(Stmt::ASSIGN, line, file);
        s->op = op; s->lhs = e; s->rhs = expr(); return s;
      }
      if (compound_base(op)) {
        p++;
        Stmt* s = unit.new_stmt(Stmt::ASSIGN, line, file);
        Expr* b = unit.new_expr(Expr::BIN, line);
        b->op = compound_base(op); b->a = e; b->b = expr();
        s->op = O_ASSIGN; s->lhs = e; s->rhs = b; return s;
      }
      if (op == O_CONSTRAIN_R || op == O_WITNESS_R) {
        p++;
        Stmt* s = unit.new_stmt(Stmt::ASSIGN, line, file);
        s->op = (op == O_CONSTRAIN_R) ? O_CONSTRAIN_L : O_WITNESS_L;
        s->rhs = e; s->lhs = expr(); return s;
      }
      if (op == O_CEQ) {
        p++;
        Stmt* s = unit.new_stmt(Stmt::CONSTR, line, file);
        s->lhs = e; s->rhs = expr(); return s;
      }
      if (op == O_INC || op == O_DEC) {
        p++;
        Stmt* s = unit.new_stmt(Stmt::INCDEC, line, file);
        s->lhs = e; s->op = (op == O_INC) ? 1 : -1; return s;
      }
    }
    err("expected statement");
  }

  Stmt* decl_signal() {
    Stmt* s = unit.new_stmt(Stmt::SIGDECL, cur().line, file);
    p++;
    s->op = 0;
    if (is_kw(kw_input)) { s->op = 1; p++; }
    else if (is_kw(kw_output)) { s->op = 2; p++; }
    if (is_op(O_LC)) { while (!eat(O_RC)) p++; }  // tags
    while (true) {
      DeclItem it; it.name = ident(); it.dims = dims();
      if (is_op(O_CONSTRAIN_L) || is_op(O_WITNESS_L)) { it.init_op = cur().op; p++; it.init = expr(); }
      s->items.push_back(it);
      if (!eat(O_COMMA)) break;
    }
    return s;
  }
  Stmt* decl_var() {
    Stmt* s = unit.new_stmt(Stmt::VARDECL, cur().line, file);
    p++;
    while (true) {
      DeclItem it; it.name = ident(); it.dims = dims();
      if (eat(O_ASSIGN)) { it.init_op = O_ASSIGN; it.init = expr(); }
      s->items.push_back(it);
      if (!eat(O_COMMA)) break;
    }
    return s;
  }
  Stmt* decl_comp() {
    Stmt* s = unit.new_stmt(Stmt::COMPDECL, cur().line, file);
    p++;
    while (is_kw(kw_parallel)) p++;
    while (true) {
      DeclItem it; it.name = ident(); it.dims = dims();
      if (eat(O_ASSIGN)) { it.init_op = O_ASSIGN; it.init = expr(); }
      s->items.push_back(it);
      if (!eat(O_COMMA)) break;
    }
    return s;
  }

  static int binprec(int op) {
    switch (op) {
      case O_LOR: return 1; case O_LAND: return 2;
      case O_EQ: case O_NE: case O_LT: case O_GT: case O_LE: case O_GE: return 3;
      case O_BOR: return 4; case O_BXOR: return 5; case O_BAND: return 6;
      case O_SHL: case O_SHR: return 7;
      case O_ADD: case O_SUB: return 8;
      case O_MUL: case O_DIV: case O_IDIV: case O_MOD: return 9;
      case O_POW: return 10;
    }
    return 0;
  }

  Expr* expr() {
    Expr* c = binexpr(1);
    if (is_op(O_QUEST)) {
      int line = cur().line; p++;
      Expr* t = unit.new_expr(Expr::TERN, line);
      t->a = c; t->b = expr(); expect(O_COLON); t->c = expr();
      return t;
    }
    return c;
  }
  Expr* binexpr(int minprec) {
    Expr* lhs = unary();
    while (cur().t == T_OP) {
      int prec = binprec(cur().op);
      if (!prec || prec < minprec) break;
      int op = cur().op, line = cur().line; p++;
      Expr* rhs = binexpr(prec + 1);
      Expr* b = unit.new_expr(Expr::BIN, line);
      b->op = op; b->a = lhs; b->b = rhs; lhs = b;
    }
    return lhs;
  }
  Expr* unary() {
    if (cur().t == T_OP && (cur().op == O_SUB || cur().op == O_LNOT || cur().op == O_BNOT)) {
      int op = cur().op, line = cur().line; p++;
      Expr* e = unary();
      if (op == O_SUB && e->k == Expr::NUM) { e->num = fr_neg(e->num); return e; }
      Expr* u = unit.new_expr(Expr::UN, line);
      u->op = (op == O_SUB) ? O_NEG : op; u->a = e;
      return u;
    }
    return postfix();
  }
  Expr* postfix() {
    Expr* e = nullptr;
    int line = cur().line;
    if (cur().t == T_NUM) { e = unit.new_expr(Expr::NUM, line); e->num = cur().num; p++; }
    else if (cur().t == T_ID) {
      if (cur().id == kw_parallel) { p++; return postfix(); }
      int name = cur().id; p++;
      if (is_op(O_LP)) {
        p++;
        e = unit.new_expr(Expr::CALL, line); e->name = name;
        while (!eat(O_RP)) { e->args.push_back(expr()); eat(O_COMMA); }
      } else { e = unit.new_expr(Expr::VAR, line); e->name = name; }
    } else if (is_op(O_LP)) { p++; e = expr(); expect(O_RP); }
    else if (is_op(O_LB)) {
      p++;
      e = unit.new_expr(Expr::ARR, line);
      while (!eat(O_RB)) { e->args.push_back(expr()); eat(O_COMMA); }
    } else err("expected expression");
    while (true) {
      if (is_op(O_LB)) {
        int l2 = cur().line; p++;
        Expr* ix = unit.new_expr(Expr::IDX, l2);
        ix->a = e; ix->b = expr(); expect(O_RB); e = ix;
      } else if (is_op(O_DOT) && peek(1).t == T_ID) {
        int l2 = cur().line; p++;
        Expr* m = unit.new_expr(Expr::MEM, l2);
        m->a = e; m->name = ident(); e = m;
      } else break;
    }
    return e;
  }
};

// realpath-ish normalisation without touching the filesystem more than needed
inline std::string norm_path(const std::string& path) {
  char buf[4096];
  if (realpath(path.c_str(), buf)) return std::string(buf);
  return path;
}
inline std::string dir_of(const std::string& path) {
  size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string(".") : path.substr(0, k);
}

inline void load_file(SourceUnit& unit, const std::string& path_in, bool is_root,
                      const std::vector<std::string>& include_dirs) {
  std::string path = norm_path(path_in);
  if (unit.seen.count(path)) return;
  unit.seen.insert(path);
  { std::ifstream probe(path); if (!probe) { unit.missing.push_back(path); return; } }
  int fid = (int)unit.files.size();
  unit.files.push_back(path);
  Parser ps(unit, path, fid);
  std::vector<std::string> incs = ps.parse_file(is_root);
  std::string d = dir_of(path);
  for (auto& inc : incs) {
    std::string cand = inc[0] == '/' ? inc : d + "/" + inc;
    { std::ifstream probe(cand);
      if (!probe) for (auto& idir : include_dirs) { std::ifstream p2(idir + "/" + inc); if (p2) { cand = idir + "/" + inc; break; } } }
    load_file(unit, cand, false, include_dirs);
  }
}

}  // namespace pzk
