// ConfigParser.h — the `.ini` reader of the reference's older API (only its header survives
// upstream: docs/ConfigParser_8h_source.html; the analytic tests read their parameters through it,
// examples/Tests.h:14-15,51,140-142). Same interface: Parser(filename), Get(section, property).
// Syntax: `[Section]`, `key = value`, comments start with ';' or '#', blank lines ignored.
#pragma once
#include <cstdlib>
#include <fstream>
#include <string>
#include <unordered_map>

#include "Exceptions.h"

struct ParserError : std::runtime_error { using std::runtime_error::runtime_error; };

struct Parser {
    using HashMap = std::unordered_map<std::string, std::unordered_map<std::string, std::string>>;

    explicit Parser(const std::string &filename) {
        std::ifstream in(filename);
        if (!in) throw ParserError("cannot open config file " + filename);
        std::string line, section;
        size_t lineno = 0;
        while (std::getline(in, line)) {
            ++lineno;
            Trim(line);
            if (line.empty() || IsComment(line)) continue;
            if (IsSection(line)) { section = line.substr(1, line.size() - 2); Trim(section); continue; }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) throw ParserError(filename + ":" + std::to_string(lineno) + ": expected key = value");
            std::string key = line.substr(0, eq), val = line.substr(eq + 1);
            const size_t c = val.find_first_of(";#");
            if (c != std::string::npos) val.erase(c);
            Trim(key); Trim(val);
            m_ini[section][key] = val;
        }
    }
    bool Has(const std::string &section, const std::string &property) const {
        const auto s = m_ini.find(section);
        return s != m_ini.end() && s->second.count(property) > 0;
    }
    double Get(const std::string &section, const std::string &property) const {
        const auto s = m_ini.find(section);
        if (s == m_ini.end()) throw ParserError("no section [" + section + "] in config");
        const auto p = s->second.find(property);
        if (p == s->second.end()) throw ParserError("no property " + property + " in section [" + section + "]");
        char *end = nullptr;
        const double v = std::strtod(p->second.c_str(), &end);
        if (end == p->second.c_str() || *end != '\0') throw ParserError("[" + section + "] " + property + " is not a number: " + p->second);
        return v;
    }
    double Get(const std::string &section, const std::string &property, double fallback) const {
        return Has(section, property) ? Get(section, property) : fallback;
    }

 private:
    static void Trim(std::string &s) {
        const char *ws = " \t\r\n";
        s.erase(0, s.find_first_not_of(ws));
        const size_t e = s.find_last_not_of(ws);
        if (e == std::string::npos) s.clear(); else s.erase(e + 1);
    }
    static bool IsComment(const std::string &line) { return line[0] == ';' || line[0] == '#'; }
    static bool IsSection(const std::string &line) { return line.front() == '[' && line.back() == ']'; }
    HashMap m_ini;
};
