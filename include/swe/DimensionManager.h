// DimensionManager.h — scaling between dimensional and the solver's non-dimensional variables
// (g = 1), as in the reference's older API (docs/DimensionManager_8h_source.html, g declared in
// docs/DimensionManager_8cpp.html). The scales h0, l0, c0 are read from section [Scales] of the
// config (the upstream key names are not recoverable; absent section => unit scales).
#pragma once
#include "ConfigParser.h"

struct DimensionError : std::runtime_error { using std::runtime_error::runtime_error; };

enum class Scales { height, length, velocity, source, time };

struct DimensionManager {
    static constexpr double g = 1.0;

    explicit DimensionManager(const Parser &parser)
        : m_h0(parser.Get("Scales", "h0", 1.0)), m_l0(parser.Get("Scales", "l0", 1.0)), m_c0(parser.Get("Scales", "c0", 1.0)) {
        if (!(m_h0 > 0) || !(m_l0 > 0) || !(m_c0 > 0)) throw DimensionError("scales h0, l0, c0 must be positive");
    }
    DimensionManager(double h0, double l0, double c0) : m_h0(h0), m_l0(l0), m_c0(c0) {}

    template <Scales T>
    double Scale(double x) const {
        switch (T) {
            case Scales::height: return x * m_h0;
            case Scales::length: return x * m_l0;
            case Scales::velocity: return x * m_c0;
            case Scales::source: return x * m_c0 / m_l0;
            case Scales::time: return x * m_l0 / m_c0;
            default: throw DimensionError("invalid dimension type in scale function");
        }
    }
    template <Scales T>
    double Unscale(double x) const {
        switch (T) {
            case Scales::height: return x / m_h0;
            case Scales::length: return x / m_l0;
            case Scales::velocity: return x / m_c0;
            case Scales::source: return x * m_l0 / m_c0;
            case Scales::time: return x * m_c0 / m_l0;
            default: throw DimensionError("invalid dimension type in unscale function");
        }
    }

 private:
    double m_h0, m_l0, m_c0;
};
