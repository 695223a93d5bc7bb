// function_expression.h — evaluator for the `Function expression` entries of a .prm file
// (deal.II Functions::ParsedFunction / FunctionParser over muparser; both are external
// dependencies of the reference). Covered is the expression language the reference's DEM
// parameter files use: numbers, the variables x, y, z, t, the constants pi / Pi, + - * / ^,
// unary minus, comparisons, && ||, `if(condition, a, b)`, `c ? a : b`, and the usual one- and
// two-argument functions. An expression is parsed once into a small tree and evaluated many
// times (per time step for solid-object velocities, per lattice site for the insertion
// acceptance function).
#pragma once

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace lethe_b200
{
  class FunctionExpression
  {
  public:
    struct Variables
    {
      double x = 0, y = 0, z = 0, t = 0;
    };

    FunctionExpression() = default;
    explicit FunctionExpression(const std::string &text)
      : source(text)
    {
      Parser p{text, 0};
      root = p.ternary();
      p.skip();
      if (p.pos != text.size())
        throw std::runtime_error("cannot parse `" + text + "` at column " + std::to_string(p.pos + 1));
    }

    bool empty() const { return !root; }
    const std::string &text() const { return source; }
    double operator()(const Variables &v) const { return root->eval(v); }
    // true when the value does not depend on x, y, z, t
    bool is_constant() const { return root && root->constant(); }

  private:
    struct Node
    {
      enum Kind { NUMBER, VARIABLE, UNARY, BINARY, CALL, IF } kind = NUMBER;
      double value = 0;
      char variable = 0;
      std::string op;
      std::vector<std::unique_ptr<Node>> args;

      bool constant() const
      {
        if (kind == VARIABLE)
          return false;
        for (const auto &a : args)
          if (!a->constant())
            return false;
        return true;
      }

      double eval(const Variables &v) const
      {
        switch (kind)
          {
            case NUMBER:
              return value;
            case VARIABLE:
              return variable == 'x' ? v.x : variable == 'y' ? v.y : variable == 'z' ? v.z : v.t;
            case UNARY:
              return op == "-" ? -args[0]->eval(v) : op == "!" ? double(args[0]->eval(v) == 0.0) : args[0]->eval(v);
            case IF:
              return args[0]->eval(v) != 0.0 ? args[1]->eval(v) : args[2]->eval(v);
            case BINARY:
              {
                const double a = args[0]->eval(v);
                if (op == "&&")
                  return double(a != 0.0 && args[1]->eval(v) != 0.0);
                if (op == "||")
                  return double(a != 0.0 || args[1]->eval(v) != 0.0);
                const double b = args[1]->eval(v);
                if (op == "+")
                  return a + b;
                if (op == "-")
                  return a - b;
                if (op == "*")
                  return a * b;
                if (op == "/")
                  return a / b;
                if (op == "^")
                  return std::pow(a, b);
                if (op == "<")
                  return double(a < b);
                if (op == ">")
                  return double(a > b);
                if (op == "<=")
                  return double(a <= b);
                if (op == ">=")
                  return double(a >= b);
                if (op == "==")
                  return double(a == b);
                return double(a != b); // "!="
              }
            case CALL:
              {
                const double a = args[0]->eval(v);
                if (args.size() == 2)
                  {
                    const double b = args[1]->eval(v);
                    if (op == "min")
                      return std::fmin(a, b);
                    if (op == "max")
                      return std::fmax(a, b);
                    if (op == "pow")
                      return std::pow(a, b);
                    return std::atan2(a, b);
                  }
                if (op == "sin")
                  return std::sin(a);
                if (op == "cos")
                  return std::cos(a);
                if (op == "tan")
                  return std::tan(a);
                if (op == "asin")
                  return std::asin(a);
                if (op == "acos")
                  return std::acos(a);
                if (op == "atan")
                  return std::atan(a);
                if (op == "sinh")
                  return std::sinh(a);
                if (op == "cosh")
                  return std::cosh(a);
                if (op == "tanh")
                  return std::tanh(a);
                if (op == "exp")
                  return std::exp(a);
                if (op == "log" || op == "ln")
                  return std::log(a);
                if (op == "log10")
                  return std::log10(a);
                if (op == "log2")
                  return std::log2(a);
                if (op == "sqrt")
                  return std::sqrt(a);
                if (op == "abs")
                  return std::fabs(a);
                if (op == "floor")
                  return std::floor(a);
                if (op == "ceil")
                  return std::ceil(a);
                if (op == "int" || op == "rint")
                  return std::rint(a);
                if (op == "sign")
                  return double((a > 0) - (a < 0));
                return std::erf(a); // "erf"
              }
          }
        return 0;
      }
    };

    struct Parser
    {
      const std::string &s;
      size_t pos;

      void skip()
      {
        while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos])))
          ++pos;
      }
      bool take(const char *token)
      {
        skip();
        const size_t n = std::char_traits<char>::length(token);
        if (s.compare(pos, n, token) != 0)
          return false;
        pos += n;
        return true;
      }
      [[noreturn]] void fail(const std::string &what) const
      {
        throw std::runtime_error("cannot parse `" + s + "`: " + what + " at column " + std::to_string(pos + 1));
      }
      static std::unique_ptr<Node> make(Node::Kind kind, const std::string &op, std::unique_ptr<Node> a, std::unique_ptr<Node> b = nullptr,
                                        std::unique_ptr<Node> c = nullptr)
      {
        auto n = std::make_unique<Node>();
        n->kind = kind;
        n->op = op;
        n->args.push_back(std::move(a));
        if (b)
          n->args.push_back(std::move(b));
        if (c)
          n->args.push_back(std::move(c));
        return n;
      }

      // precedence, lowest first: ?: , || , && , comparisons , + - , * / , unary - , ^
      std::unique_ptr<Node> ternary()
      {
        auto c = logical_or();
        if (!take("?"))
          return c;
        auto a = ternary();
        if (!take(":"))
          fail("expected `:`");
        auto b = ternary();
        return make(Node::IF, "?", std::move(c), std::move(a), std::move(b));
      }
      std::unique_ptr<Node> logical_or()
      {
        auto a = logical_and();
        while (take("||") || (peek_single('|') && take("|")))
          a = make(Node::BINARY, "||", std::move(a), logical_and());
        return a;
      }
      std::unique_ptr<Node> logical_and()
      {
        auto a = comparison();
        while (take("&&") || (peek_single('&') && take("&")))
          a = make(Node::BINARY, "&&", std::move(a), comparison());
        return a;
      }
      bool peek_single(char c)
      {
        skip();
        return pos < s.size() && s[pos] == c;
      }
      std::unique_ptr<Node> comparison()
      {
        auto a = sum();
        for (;;)
          {
            const char *found = nullptr;
            for (const char *op : {"<=", ">=", "==", "!=", "<", ">"})
              if (take(op))
                {
                  found = op;
                  break;
                }
            if (!found)
              return a;
            a = make(Node::BINARY, found, std::move(a), sum());
          }
      }
      std::unique_ptr<Node> sum()
      {
        auto a = product();
        for (;;)
          {
            if (take("+"))
              a = make(Node::BINARY, "+", std::move(a), product());
            else if (take("-"))
              a = make(Node::BINARY, "-", std::move(a), product());
            else
              return a;
          }
      }
      std::unique_ptr<Node> product()
      {
        auto a = unary();
        for (;;)
          {
            if (take("*"))
              a = make(Node::BINARY, "*", std::move(a), unary());
            else if (take("/"))
              a = make(Node::BINARY, "/", std::move(a), unary());
            else
              return a;
          }
      }
      std::unique_ptr<Node> unary()
      {
        if (take("-"))
          return make(Node::UNARY, "-", unary());
        if (take("+"))
          return unary();
        if (take("!"))
          return make(Node::UNARY, "!", unary());
        return power();
      }
      std::unique_ptr<Node> power()
      {
        auto base = primary();
        if (take("^"))
          return make(Node::BINARY, "^", std::move(base), unary()); // right associative
        return base;
      }
      std::unique_ptr<Node> primary()
      {
        skip();
        if (pos >= s.size())
          fail("unexpected end");
        if (take("("))
          {
            auto e = ternary();
            if (!take(")"))
              fail("expected `)`");
            return e;
          }
        const unsigned char c = static_cast<unsigned char>(s[pos]);
        if (std::isdigit(c) || c == '.')
          {
            char *end = nullptr;
            const double v = std::strtod(s.c_str() + pos, &end);
            pos = size_t(end - s.c_str());
            auto n = std::make_unique<Node>();
            n->value = v;
            return n;
          }
        if (std::isalpha(c) || c == '_')
          {
            size_t e = pos;
            while (e < s.size() && (std::isalnum(static_cast<unsigned char>(s[e])) || s[e] == '_'))
              ++e;
            const std::string name = s.substr(pos, e - pos);
            pos = e;
            if (take("("))
              {
                std::vector<std::unique_ptr<Node>> args;
                args.push_back(ternary());
                while (take(","))
                  args.push_back(ternary());
                if (!take(")"))
                  fail("expected `)`");
                if (name == "if")
                  {
                    if (args.size() != 3)
                      fail("if() takes 3 arguments");
                    return make(Node::IF, "if", std::move(args[0]), std::move(args[1]), std::move(args[2]));
                  }
                static const char *one[] = {"sin",  "cos", "tan",  "asin", "acos", "atan",  "sinh", "cosh", "tanh", "exp", "log", "ln",
                                            "log10", "log2", "sqrt", "abs",  "floor", "ceil", "int",  "rint", "sign", "erf"};
                static const char *two[] = {"min", "max", "pow", "atan2"};
                bool known = false;
                for (const char *f : one)
                  known = known || (name == f && args.size() == 1);
                for (const char *f : two)
                  known = known || (name == f && args.size() == 2);
                if (!known)
                  fail("unknown function `" + name + "` with " + std::to_string(args.size()) + " argument(s)");
                auto n = std::make_unique<Node>();
                n->kind = Node::CALL;
                n->op = name;
                n->args = std::move(args);
                return n;
              }
            auto n = std::make_unique<Node>();
            if (name == "pi" || name == "Pi" || name == "PI")
              n->value = M_PI;
            else if (name == "x" || name == "y" || name == "z" || name == "t")
              {
                n->kind = Node::VARIABLE;
                n->variable = name[0];
              }
            else
              fail("unknown name `" + name + "`");
            return n;
          }
        fail("unexpected character");
      }
    };

    std::string source;
    std::shared_ptr<Node> root; // shared: the tree is immutable once parsed
  };
} // namespace lethe_b200
