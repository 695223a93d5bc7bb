"""CPU oracle — TEST INFRASTRUCTURE ONLY (see dem_oracle.cpp)."""
