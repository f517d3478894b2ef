# Builds the in-tree C-ABI library (sm_100a only) and the C oracle helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVBASE := $(ARCH) -O3 -lineinfo -std=c++17 -Xptxas -v
# front end: bit-exact parity with OpenCV's non-FMA float arithmetic -> no contraction anywhere
NVFLAGS := $(NVBASE) -Xcompiler -fPIC,-O3,-ffp-contract=off --fmad=false
# back end: FP64 linear algebra, FMA allowed
NVFLAGS_BE := $(NVBASE) -Xcompiler -fPIC,-O3
SRC_CU := $(wildcard larvio_b200/csrc/*.cu)
SRC_CPP := $(wildcard larvio_b200/csrc/*.cpp)
OBJ := $(SRC_CU:.cu=.o) $(SRC_CPP:.cpp=.o)
LIB := larvio_b200/lib/liblarvio_b200.so

IOLIB := larvio_b200/lib/liblarvio_io.so
REPLAY := larvio_b200/bin/larvio_replay
SHIMDEMO := larvio_b200/bin/larvio_shim_demo

ORACLE_C := oracle/_build/liboracle_backend.so oracle/_build/liboracle_orb.so

all: $(LIB) $(IOLIB) $(REPLAY) $(SHIMDEMO) $(ORACLE_C)

# the compiled CPU oracle of processFeatures (test infrastructure / bench CPU legs only; never linked into the product)
oracle/_build/liboracle_backend.so: oracle/backend_c.cpp
	mkdir -p oracle/_build
	g++ -O3 -march=x86-64-v3 -std=c++17 -shared -fPIC -o $@ $<

oracle/_build/liboracle_orb.so: oracle/orb_c.cpp
	mkdir -p oracle/_build
	g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -shared -fPIC -o $@ $<

# host-side on-disk formats (PNG/CSV readers, links zlib) and the C++ batched replay driver (app/larvioMain.cpp's role)
$(IOLIB): larvio_b200/host/lvb_io.cpp include/larvio_b200.h
	mkdir -p larvio_b200/lib
	g++ -O2 -std=c++17 -fPIC -shared -o $@ $< -lz

$(REPLAY): larvio_b200/host/replay_main.cpp $(LIB) $(IOLIB) include/larvio_b200.h
	mkdir -p larvio_b200/bin
	g++ -O2 -std=c++17 -o $@ $< -Llarvio_b200/lib -llarvio_b200 -llarvio_io -Wl,-rpath,'$$ORIGIN/../lib'

# the drop-in facade (larvio_shim.hpp: the reference's two classes) linked and run as the reference's own main loop
$(SHIMDEMO): larvio_b200/host/shim_main.cpp larvio_b200/host/larvio_shim.hpp $(LIB) $(IOLIB) include/larvio_b200.h
	mkdir -p larvio_b200/bin
	g++ -O2 -std=c++17 -o $@ $< -Llarvio_b200/lib -llarvio_b200 -llarvio_io -Wl,-rpath,'$$ORIGIN/../lib'

larvio_b200/csrc/be_%.o: larvio_b200/csrc/be_%.cu larvio_b200/csrc/*.h larvio_b200/csrc/*.cuh larvio_b200/csrc/*.inc include/larvio_b200.h
	$(NVCC) $(NVFLAGS_BE) -c $< -o $@ 2> $@.log || (cat $@.log; false)

larvio_b200/csrc/%.o: larvio_b200/csrc/%.cu larvio_b200/csrc/*.h larvio_b200/csrc/*.cuh larvio_b200/csrc/*.inc include/larvio_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.log || (cat $@.log; false)

larvio_b200/csrc/%.o: larvio_b200/csrc/%.cpp include/larvio_b200.h
	g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p larvio_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -lpthread

clean:
	rm -f larvio_b200/csrc/*.o larvio_b200/csrc/*.o.log $(LIB) $(IOLIB) $(REPLAY) $(SHIMDEMO) $(ORACLE_C)

# The reference's own filter (src/larvio.cpp + the static initialiser, compiled UNMODIFIED from where they lie) against the
# stand-in headers of oracle/ref_shim/ (Eigen / boost / OpenCV-core subsets written for this purpose; none of them is in the
# image).  Test infrastructure: tests/golden/make_ref_golden.py runs it to produce the golden vectors that pin the oracle
# and the CUDA back end.  Only buildable where /root/reference exists; outputs only into oracle/_ref/ (git-ignored).
REF_SRC := /root/reference
REF_FLAGS := -O2 -std=c++17 -w -Ioracle/ref_shim -I$(REF_SRC)/include
ref: oracle/_ref/larvio_ref
oracle/_ref/larvio_ref: oracle/ref_driver.cpp oracle/ref_shim/Eigen/Dense oracle/ref_shim/opencv2/lvb_cv.hpp oracle/ref_shim/boost/math/distributions/chi_squared.hpp oracle/ref_shim/Initializer/DynamicInitializer.h
	mkdir -p oracle/_ref
	g++ $(REF_FLAGS) -o $@ oracle/ref_driver.cpp $(REF_SRC)/src/larvio.cpp $(REF_SRC)/src/StaticInitializer.cpp $(REF_SRC)/src/FlexibleInitializer.cpp

# The reference's own FRONT END (src/image_processor.cpp + src/ORBDescriptor.cpp, compiled UNMODIFIED) against the same stand-in
# headers; their OpenCV functions are executed by cv2 through oracle/cv_server.py (no OpenCV C++ headers in the image).
ref_fe: oracle/_ref/larvio_ref_fe
oracle/_ref/larvio_ref_fe: oracle/ref_fe_driver.cpp oracle/ref_shim/opencv2/lvb_cv.hpp oracle/ref_shim/Eigen/Dense
	mkdir -p oracle/_ref
	g++ $(REF_FLAGS) -ffp-contract=off -o $@ oracle/ref_fe_driver.cpp $(REF_SRC)/src/image_processor.cpp $(REF_SRC)/src/ORBDescriptor.cpp

# ... and the whole per-frame pipeline: front end + filter + static initialiser behind the loop of app/larvioMain.cpp:87-117
ref_main: oracle/_ref/larvio_ref_main
oracle/_ref/larvio_ref_main: oracle/ref_main_driver.cpp oracle/ref_shim/opencv2/lvb_cv.hpp oracle/ref_shim/Eigen/Dense
	mkdir -p oracle/_ref
	g++ $(REF_FLAGS) -ffp-contract=off -o $@ oracle/ref_main_driver.cpp $(REF_SRC)/src/image_processor.cpp $(REF_SRC)/src/ORBDescriptor.cpp $(REF_SRC)/src/larvio.cpp $(REF_SRC)/src/StaticInitializer.cpp $(REF_SRC)/src/FlexibleInitializer.cpp
