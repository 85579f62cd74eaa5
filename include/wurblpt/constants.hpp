/* constants.hpp -- physical constants and debug tolerances of the WurblPT API
 * (mirror of reference libwurblpt/constants.hpp:30-42). */
#pragma once

namespace WurblPT {

constexpr float refractiveIndexOfVacuum = 1.0f;
constexpr float refractiveIndexOfAir = 1.00028f;
constexpr double speedOfLight = 299792458;          // m/s
constexpr double elementaryCharge = 0.1602176565;   // attocoulomb
constexpr float hc = 1.98644582;                    // Planck constant * speed of light, 1e-25 m^3 kg / s^2
constexpr float cosOrthoAngleTolerance = 0.0003f;
constexpr float dirSquaredLengthTolerance = 0.00001f;

}
