// toml_fwd.hpp — stand-in for toml11's forward-declaration header (see toml.hpp)
#ifndef OCM_REF_SHIM_TOML_FWD_HPP
#define OCM_REF_SHIM_TOML_FWD_HPP
#include "toml.hpp"
#endif
