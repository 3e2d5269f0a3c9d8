# Shared recipe: compile the reference's own translation units, where they lie under $(REF), into $(OUT).
# No reference source is copied; the only generated inputs are artisoptions.h (preset + the config's
# overrides, written by tools/gen_inputs.py) and version.h. The reference's own build system is not used.
#
#   make -f oracle/ref_build.mk CONFIG=<tools/configs.py name> FLAVOR=parity|fast KIND=oracle|dropin OUT=<dir>
#
#   FLAVOR=parity : -DGPU_ON (per-packet RNG state, all cell caches precomputed) with the reference's
#                   REPRODUCIBLE flags (Makefile:111-114): -O2 -ffp-contract=off, no fast-math
#   FLAVOR=fast   : the reference's production flags (Makefile:236-252): -O3 -march=native -ffast-math ...
#   KIND=oracle   : update_packets.cc of the reference is embedded (renamed) next to the snapshot hooks
#   KIND=dropin   : update_packets() is provided by the B200 binding only (reference update_packets.cc not linked)
REF ?= /root/reference
REPO ?= $(abspath $(dir $(lastword $(MAKEFILE_LIST)))/..)
CONFIG ?= classic_toy
FLAVOR ?= parity
KIND ?= oracle
OUT ?= $(REPO)/oracle/_ref/$(CONFIG)/$(FLAVOR)
CXX ?= g++

ACCESS := $(REPO)/integration/ref_access
INCLUDES := -I$(OUT) -I$(REPO)/oracle/shim -I$(REF) -I$(REPO)/integration -I$(ACCESS) -I$(REPO)/include -isystem $(REF)/third_party
CXXFLAGS := -std=c++23 -w $(INCLUDES)
ifeq ($(FLAVOR),parity)
  CXXFLAGS += -O2 -DREPRODUCIBLE=true -ffp-contract=off -DEIGEN_DONT_VECTORIZE -DGPU_ON
else
  CXXFLAGS += -O3 -march=native -ffast-math -funsafe-math-optimizations -fno-finite-math-only
endif
ifeq ($(KIND),oracle)
  CXXFLAGS += -DARTISB200_WITH_REFERENCE
  BIN := sn3d_ref
else
  BIN := sn3d_b200
endif

# reference TUs compiled directly, except those wrapped by integration/ref_access (which #include them)
WRAPPED := grid ratecoeff kpkt radfield stats nonthermal rpkt gammapkt spectrum_lightcurve
PLAIN := $(filter-out $(WRAPPED) update_packets exspec unittests sn3d,$(basename $(notdir $(wildcard $(REF)/*.cc))))
OBJS := $(addprefix $(OUT)/,$(addsuffix .o,$(PLAIN) sn3d $(addprefix ref_,$(WRAPPED)) update_packets_b200))

all: $(OUT)/$(BIN)

$(OUT)/artisoptions.h:
	mkdir -p $(OUT)
	python3 $(REPO)/tools/gen_inputs.py $(CONFIG) $(OUT)/inputs --reference $(REF) --options-out $@
	echo 'constexpr const char* GIT_VERSION="artis_b200-refbuild"; constexpr const char* GIT_BRANCH="$(KIND)-$(FLAVOR)"; constexpr const char* GIT_STATUS="";' > $(OUT)/version.h

$(OUT)/%.o: $(REF)/%.cc $(OUT)/artisoptions.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OUT)/ref_%.o: $(ACCESS)/ref_%.cc $(OUT)/artisoptions.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OUT)/update_packets_b200.o: $(REPO)/integration/update_packets_b200.cc $(OUT)/artisoptions.h $(REPO)/include/artis_b200.h $(REPO)/include/artis_b200_options.h $(REPO)/integration/b200_snapshot.h $(ACCESS)/b200_access.h
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OUT)/$(BIN): $(OBJS)
	$(CXX) -o $@ $^ -ldl

# exspec (the reference's post-processor) for spectra comparisons
EXSPEC_OBJS := $(addprefix $(OUT)/,$(addsuffix .o,$(PLAIN) exspec $(addprefix ref_,$(WRAPPED)) update_packets_b200))
$(OUT)/exspec: $(EXSPEC_OBJS)
	$(CXX) -o $@ $^ -ldl

# object files are intermediates: a missing .o does not rebuild an up-to-date binary (they are deleted to keep the
# snapshot that travels to the GPU box small)
.SECONDARY: $(OBJS) $(EXSPEC_OBJS)

.PHONY: all
