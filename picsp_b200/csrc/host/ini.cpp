// INI reader + the reference's normalisation, sanity gates and banner (src/main.cpp:240-331).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "host.hpp"

namespace picsp_host {

static std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
static std::string lower(std::string s) {
    for (auto &ch : s) ch = (char)std::tolower((unsigned char)ch);
    return s;
}

// One logical line -> (section | key,value).  Rules of iniparser_line(): blank and '#'/';' lines are
// skipped; "[name]" opens a section (stripped, lower-cased); "key = value" where the value is
// "double quoted", 'single quoted', or everything up to the first ';' or '#'; key stripped and
// lower-cased, value stripped; `key=`, `key=;`, `key=#` give an empty value.
bool IniFile::load(const std::string &path, std::string *err) {
    std::ifstream in(path);
    if (!in) { if (err) *err = "cannot parse file: " + path; return false; }
    std::string raw, line, section;
    while (std::getline(in, raw)) {
        // a trailing backslash joins the next line (iniparser_load multi-line support)
        line += raw;
        std::string t = strip(line);
        if (!t.empty() && t.back() == '\\') { line = t.substr(0, t.size() - 1); continue; }
        line.clear();
        if (t.empty() || t[0] == '#' || t[0] == ';') continue;
        if (t.front() == '[' && t.back() == ']') {
            size_t close = t.find(']');
            section = lower(strip(t.substr(1, close - 1)));
            kv_[section] = "";
            continue;
        }
        size_t eq = t.find('=');
        if (eq == std::string::npos || eq == 0) { if (err) *err = "syntax error in " + path + ": " + t; return false; }
        std::string key = lower(strip(t.substr(0, eq)));
        std::string rest = strip(t.substr(eq + 1));
        std::string val;
        // iniparser_line() tries a double-quoted form, then a single-quoted one, then "everything up to ; or #"
        // (iniparser.c:583-585).  A quoted value needs at least one character that is not the quote, and sscanf does
        // not insist on the closing quote; an empty pair of quotes falls through to the third form and is then mapped
        // to the empty string (:594-596), while three or more quotes in a row stay as they are.
        if (rest.size() >= 2 && rest[0] == '"' && rest[1] != '"') {
            const size_t close = rest.find('"', 1);
            val = strip(rest.substr(1, close == std::string::npos ? std::string::npos : close - 1));
        } else if (rest.size() >= 2 && rest[0] == '\'' && rest[1] != '\'') {
            const size_t close = rest.find('\'', 1);
            val = strip(rest.substr(1, close == std::string::npos ? std::string::npos : close - 1));
        } else {
            size_t cut = rest.find_first_of(";#");
            val = strip(cut == std::string::npos ? rest : rest.substr(0, cut));
            if (val == "\"\"" || val == "''") val.clear();
        }
        kv_[section.empty() ? key : section + ":" + key] = val;
    }
    return true;
}
bool IniFile::has(const std::string &key) const { return kv_.count(lower(key)) != 0; }
int IniFile::get_int(const std::string &key, int notfound) const {
    auto it = kv_.find(lower(key));
    if (it == kv_.end()) return notfound;
    return (int)std::strtol(it->second.c_str(), nullptr, 0);
}
double IniFile::get_double(const std::string &key, double notfound) const {
    auto it = kv_.find(lower(key));
    if (it == kv_.end()) return notfound;
    return std::atof(it->second.c_str());
}

int parse_run_config(const std::string &path, picsp_run_config &c, bool banner, std::string *err) {
    IniFile ini;
    if (!ini.load(path, err)) return PICSP_ERR_INVALID;
    // src/main.cpp:58-61
    const double EPS_un = 8.85418782E-12, K = 1.38065E-23, EV_TO_K = 11604.52;
    c.nTimeSteps = ini.get_int("time:nTimeSteps", -1);
    const double timeStep_unorm = ini.get_double("time:timeStep", -1.0);
    const double stepSize_unorm = ini.get_double("grid:stepSize", -1.0);
    c.numxCells = ini.get_int("grid:numxCells", -1);
    c.numyCells = ini.get_int("grid:numyCells", -1);
    c.nParticlesI = ini.get_int("population:nParticlesI", -1);
    c.nParticlesE = ini.get_int("population:nParticlesE", -1);
    const double massI_unorm = ini.get_double("population:massI", -1.0);
    const double massE_unorm = ini.get_double("population:massE", -1.0);
    const double chargeE_unorm = ini.get_double("population:chargeE", -1.0);
    const double density_unorm = ini.get_double("population:density", -1.0);
    const double vthE_unorm = ini.get_double("population:vthE", -1.0);
    const double vthI_unorm = ini.get_double("population:vthI", -1.0);
    const double driftE_unorm = ini.get_double("population:driftE", -1.0);
    const double driftI_unorm = ini.get_double("population:driftI", -1.0);
    c.dumpPeriod = ini.get_int("diagnostics:dumpPeriod", -1);
    c.solverType = (short)ini.get_int("solver:solverType", -1);     // `short int` in the reference (main.cpp:66-67)
    c.loadType = (short)ini.get_int("population:loadType", -1);

    // normalisation, src/main.cpp:279-294
    c.omega_pe = std::sqrt((chargeE_unorm * chargeE_unorm * density_unorm) / (massE_unorm * EPS_un));
    c.Lambda_D = std::sqrt((EPS_un * K * vthE_unorm * EV_TO_K) / (density_unorm * chargeE_unorm * chargeE_unorm));
    c.chargeE = chargeE_unorm / chargeE_unorm;
    c.massI = massI_unorm / massE_unorm;
    c.massE = massE_unorm / massE_unorm;
    c.driftE = driftE_unorm / vthE_unorm;
    c.driftI = driftI_unorm / vthE_unorm;
    c.density = density_unorm / density_unorm;
    c.timeStep = timeStep_unorm * c.omega_pe;
    c.stepSize = stepSize_unorm / c.Lambda_D;
    c.vthE = vthE_unorm / vthE_unorm;
    c.vthI = vthI_unorm / vthE_unorm;
    c.ion_spwt = (c.density * c.numxCells * c.numyCells * c.stepSize * c.stepSize) / (c.nParticlesI);
    c.electron_spwt = (c.density * c.numxCells * c.numyCells * c.stepSize * c.stepSize) / (c.nParticlesE);

    std::ostream &o = std::cout;
    if (banner) {
        o << "********** IMPORTANT PLASMA QUANTITIES ***********" << '\n';
        o << "omega_pe: " << c.omega_pe << std::endl;
        o << "Lambda_D: " << c.Lambda_D << std::endl;
        o << "*************** Input Sanity Check ***************" << '\n';
    }
    bool ok = true;
    std::string why;
    if (c.stepSize >= 1) { why += "ERROR, stepSize is bigger than Debye length.\n"; ok = false; }
    if (c.timeStep > 0.01) {
        char buf[160];
        std::snprintf(buf, sizeof buf, "ERROR, timeStep is too big. The recommended value: <%g s\n", 0.01 / c.omega_pe);
        why += buf; ok = false;
    }
    if (c.solverType != 1 && c.solverType != 2) {
        why += "ERROR, Wrong Solver Type. The recommended value: 1 or 2\nsolverType: " + std::to_string(c.solverType) + "\n";
        ok = false;
    }
    if (c.loadType != 1 && c.loadType != 2) {
        why += "ERROR, Wrong Load Type. The recommended value: 1 or 2\nloadType: " + std::to_string(c.loadType) + "\n";
        ok = false;
    }
    if (banner) {
        o << why;
        if (ok) o << "STATUS, Input parameters are compatible." << std::endl;
        else o << "ERROR, Input parameters are incompatible." << std::endl;
    }
    if (!ok) { if (err) *err = why; return PICSP_ERR_INVALID; }   // the reference exit(EXIT_FAILURE)s here
    return PICSP_OK;
}

}  // namespace picsp_host
