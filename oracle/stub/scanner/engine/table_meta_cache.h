// Stand-in for the reference's scanner/engine/table_meta_cache.h (storehouse-backed; not needed by sampler.cpp).
#pragma once
