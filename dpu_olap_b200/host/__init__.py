"""C++ host layer: the *Gpu siblings of the reference's *Dpu operator classes over Arrow C++,
the *Native (Acero) classes restated for Arrow 24, the generator, and the test / benchmark
drivers. Built by build_host.build() into dpu_olap_b200/host/_build/."""
