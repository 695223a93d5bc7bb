// prm_file.h — reader for deal.II ParameterHandler `.prm` files:
//   `subsection <name>` … `end`, `set <key> = <value>`, `#` comments.
// Keys and subsection names are kept verbatim, so the host code below looks entries up by
// the exact spellings the reference declares (source/core/parameters_lagrangian.cc,
// source/core/parameters.cc).
#pragma once

#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace lethe_b200
{
  class PrmSection
  {
  public:
    std::map<std::string, std::string> values;
    std::map<std::string, std::unique_ptr<PrmSection>> children;

    const PrmSection &sub(const std::string &name) const
    {
      static const PrmSection empty;
      auto it = children.find(name);
      return it == children.end() ? empty : *it->second;
    }
    bool has(const std::string &key) const { return values.count(key) != 0; }
    bool has_sub(const std::string &name) const { return children.count(name) != 0; }
    std::string get(const std::string &key, const std::string &fallback) const
    {
      auto it = values.find(key);
      return it == values.end() ? fallback : it->second;
    }
    double get_double(const std::string &key, double fallback) const
    {
      return has(key) ? std::stod(values.at(key)) : fallback;
    }
    long get_int(const std::string &key, long fallback) const { return has(key) ? std::stol(values.at(key)) : fallback; }
    bool get_bool(const std::string &key, bool fallback) const
    {
      if (!has(key))
        return fallback;
      const std::string v = lower(values.at(key));
      return v == "true" || v == "1" || v == "yes" || v == "on";
    }
    // comma (or `sep`) separated list of doubles
    std::vector<double> get_list(const std::string &key, char sep = ',') const { return split_doubles(get(key, ""), sep); }

    static std::string trim(const std::string &s)
    {
      const auto a = s.find_first_not_of(" \t\r\n");
      if (a == std::string::npos)
        return "";
      const auto b = s.find_last_not_of(" \t\r\n");
      return s.substr(a, b - a + 1);
    }
    static std::string lower(std::string s)
    {
      for (auto &c : s)
        c = char(std::tolower(static_cast<unsigned char>(c)));
      return s;
    }
    static std::vector<std::string> split(const std::string &s, char sep)
    {
      std::vector<std::string> out;
      std::string item;
      std::istringstream in(s);
      while (std::getline(in, item, sep))
        out.push_back(trim(item));
      return out;
    }
    static std::vector<double> split_doubles(const std::string &s, char sep)
    {
      std::vector<double> out;
      for (const auto &t : split(s, sep))
        if (!t.empty())
          out.push_back(std::stod(t));
      return out;
    }
  };

  inline std::unique_ptr<PrmSection> parse_prm(std::istream &in)
  {
    auto root = std::make_unique<PrmSection>();
    std::vector<PrmSection *> stack{root.get()};
    std::string raw;
    int line_no = 0;
    while (std::getline(in, raw))
      {
        ++line_no;
        const auto hash = raw.find('#');
        const std::string line = PrmSection::trim(hash == std::string::npos ? raw : raw.substr(0, hash));
        if (line.empty())
          continue;
        if (line.rfind("subsection", 0) == 0 && (line.size() == 10 || std::isspace(static_cast<unsigned char>(line[10]))))
          {
            const std::string name = PrmSection::trim(line.substr(10));
            auto &slot = stack.back()->children[name];
            if (!slot)
              slot = std::make_unique<PrmSection>();
            stack.push_back(slot.get());
          }
        else if (line == "end")
          {
            if (stack.size() == 1)
              throw std::runtime_error("prm line " + std::to_string(line_no) + ": unbalanced `end`");
            stack.pop_back();
          }
        else if (line.rfind("set", 0) == 0 && line.size() > 3 && std::isspace(static_cast<unsigned char>(line[3])))
          {
            const auto eq = line.find('=');
            if (eq == std::string::npos)
              throw std::runtime_error("prm line " + std::to_string(line_no) + ": `set` without `=`");
            stack.back()->values[PrmSection::trim(line.substr(3, eq - 3))] = PrmSection::trim(line.substr(eq + 1));
          }
        else
          throw std::runtime_error("prm line " + std::to_string(line_no) + ": cannot parse `" + raw + "`");
      }
    if (stack.size() != 1)
      throw std::runtime_error("prm: unterminated subsection");
    return root;
  }
} // namespace lethe_b200
