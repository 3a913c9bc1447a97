// main.cpp -- `huffman-codec` command line on top of the B200 library: same options, stderr
// text and exit codes as the reference CLI (src/main.cpp:22-35, 152-221), so a
// script that calls the reference binary can call this one.  Extensions that do not change any
// reference behaviour (the batch front-end of SURVEY.md 8(f)1; the .out format of a file never changes):
//   -i may be given several times, -L LIST names a file with one input path per line: all files go
//      through ONE batched GPU call; outputs are OFILE, OFILE.1, OFILE.2 ...
//   -X INDEX  batch container: compression writes all .out files back to back (starts aligned to 16
//      bytes) into OFILE and one line "<offset> <size> <input path>" per file into INDEX;
//      decompression (-d -i CONTAINER -X INDEX) splits the container by that index again.
// A -w argument that is not a number ends with the message and exit code of an invalid width (4)
// instead of the reference's uncaught exception.
#include <unistd.h>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "hc_host.hpp"

using namespace std;

static const string HELP_MESSAGE =
    "USAGE:\n"
    "  huffman-codec [-cm] -i IFILE [-o OFILE]\n"
    "  huffman-codec [-cm] -a [-w WIDTH] -i IFILE [-o OFILE]\n"
    "  huffman-codec -d -i IFILE [-o OFILE] | -h\n"
    "\n"
    "OPTION:\n"
    "  -c/-d  perform compression/decompression\n"
    "  -m     use differential model for preprocessing\n"
    "  -a     use adaptive block RLE (default: RLE)\n"
    "  -w     width of 2D data (default: 512)\n"
    "  -i     input file path\n"
    "  -o     output file path (default: b.out)\n"
    "  -h     show this help\n";

static void cerrh(const char *s) { cerr << s << "try 'huffman-codec -h' for more information\n"; }

int main(int argc, char *argv[])
{
    bool useCompr = true, useDiffModel = false, useAdaptRLE = false;
    vector<string> ifps;
    string ofp = "b.out", listp, indexp;
    uint64_t matrixWidth = 512;

    int opt;
    while ((opt = getopt(argc, argv, ":cdmai:o:w:hL:X:")) != -1) {
        switch (opt) {
        case 'c': useCompr = true; break;
        case 'd': useCompr = false; break;
        case 'm': useDiffModel = true; break;
        case 'a': useAdaptRLE = true; break;
        case 'i': ifps.push_back(optarg); break;
        case 'o': ofp = optarg; break;
        case 'w':
            try { matrixWidth = stoull(optarg); }
            catch (const exception &) { cerrh("ERROR: invalid 2D data width\n"); return 4; }
            break;
        case 'L': listp = optarg; break;
        case 'X': indexp = optarg; break;
        case 'h': cout << HELP_MESSAGE; return 0;
        case ':': cerrh("ERROR: missing additional argument\n"); return 1;
        case '?': cerrh("ERROR: unrecognized option used\n"); return 2;
        }
    }
    if (!listp.empty()) {
        ifstream lf(listp);
        if (lf.fail()) { cerr << "ERROR: given input file does not exist\n"; return 5; }
        for (string ln; getline(lf, ln);)
            if (!ln.empty()) ifps.push_back(ln);
    }
    if (ifps.empty() || ifps.back().empty()) { cerrh("ERROR: no input file path provided\n"); return 3; }
    if (useCompr && matrixWidth == 0) { cerrh("ERROR: invalid 2D data width\n"); return 4; }

    vector<vector<uint8_t>> inputs;
    for (auto &p : ifps) {
        ifstream ifs(p, ios::in | ios::binary);
        if (ifs.fail()) { cerr << "ERROR: given input file does not exist\n"; return 5; }
        inputs.emplace_back(istreambuf_iterator<char>(ifs), istreambuf_iterator<char>());
    }
    vector<string> names = ifps;
    if (!useCompr && !indexp.empty()) {
        // a batch container: cut the (single) input by its index
        ifstream xf(indexp);
        if (xf.fail() || inputs.size() != 1) { cerr << "ERROR: given input file does not exist\n"; return 5; }
        const vector<uint8_t> all = inputs[0];
        inputs.clear();
        names.clear();
        uint64_t off, len;
        string name;
        while (xf >> off >> len && getline(xf, name)) {
            if (off > all.size() || len > all.size() - off) { cerr << "ERROR: invalid or missing Huffman coding header\n"; return 8; }
            inputs.emplace_back(all.begin() + (ptrdiff_t)off, all.begin() + (ptrdiff_t)(off + len));
            names.push_back(name);
        }
    }

    vector<vector<uint8_t>> outputs;
    vector<int> status;
    try {
        hc::Codec codec(0);
        if (useCompr) outputs = codec.huffCompressBatch(inputs, useDiffModel, useAdaptRLE, matrixWidth, status);
        else outputs = codec.huffDecompressBatch(inputs, status);
    } catch (const hc::CodecError &e) {
        cerr << "ERROR: " << e.what() << "\n";
        return 70;                                             // no reference analogue: GPU/library failure
    } catch (const exception &e) {
        cerr << "ERROR: " << e.what() << "\n";                 // e.g. std::bad_alloc
        return 70;
    }
    for (size_t i = 0; i < outputs.size(); i++) {
        if (status[i] != 0) {
            cerr << "ERROR: " << hc::statusMessage(status[i]) << "\n";
            return status[i];
        }
    }
    if (useCompr && !indexp.empty()) {
        // batch container + offsets index
        ofstream ofs(ofp, ios::out | ios::binary), xf(indexp);
        uint64_t pos = 0;
        for (auto &o : outputs) pos += (o.size() + 15) / 16 * 16;
        cerr << "writing " << pos << " bytes to " << ofp << "\n";
        if (ofs.fail() || xf.fail()) { cerr << "ERROR: cannot write to " << ofp << " output file\n"; return 7; }
        pos = 0;
        static const char zeros[16] = {0};
        for (size_t i = 0; i < outputs.size(); i++) {
            xf << pos << " " << outputs[i].size() << " " << names[i] << "\n";
            ofs.write((const char *)outputs[i].data(), (streamsize)outputs[i].size());
            const uint64_t padded = (outputs[i].size() + 15) / 16 * 16;
            ofs.write(zeros, (streamsize)(padded - outputs[i].size()));
            pos += padded;
        }
        return 0;
    }
    for (size_t i = 0; i < outputs.size(); i++) {
        string path = i == 0 ? ofp : ofp + "." + to_string(i);
        cerr << "writing " << outputs[i].size() << " bytes to " << path << "\n";
        ofstream ofs(path, ios::out | ios::binary);
        if (ofs.fail()) { cerr << "ERROR: cannot write to " << path << " output file\n"; return 7; }
        ofs.write((const char *)outputs[i].data(), (streamsize)outputs[i].size());
    }
    return 0;
}
