// main.cpp -- `huffman-codec` command line on top of the B200 library: same options, stderr
// text and exit codes as the reference CLI (src/main.cpp:22-35, 152-221), so a
// script that calls the reference binary can call this one.  One extension that does not change
// any reference behaviour: `-i` may be given several times (outputs OFILE, OFILE.1, OFILE.2 ...),
// all files going through ONE batched GPU call.
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
    string ofp = "b.out";
    uint64_t matrixWidth = 512;

    int opt;
    while ((opt = getopt(argc, argv, ":cdmai:o:w:h")) != -1) {
        switch (opt) {
        case 'c': useCompr = true; break;
        case 'd': useCompr = false; break;
        case 'm': useDiffModel = true; break;
        case 'a': useAdaptRLE = true; break;
        case 'i': ifps.push_back(optarg); break;
        case 'o': ofp = optarg; break;
        case 'w': matrixWidth = stoull(optarg); break;     // non-numeric: uncaught exception, as in the reference
        case 'h': cout << HELP_MESSAGE; return 0;
        case ':': cerrh("ERROR: missing additional argument\n"); return 1;
        case '?': cerrh("ERROR: unrecognized option used\n"); return 2;
        }
    }
    if (ifps.empty() || ifps.back().empty()) { cerrh("ERROR: no input file path provided\n"); return 3; }
    if (useCompr && matrixWidth == 0) { cerrh("ERROR: invalid 2D data width\n"); return 4; }

    vector<vector<uint8_t>> inputs;
    for (auto &p : ifps) {
        ifstream ifs(p, ios::in | ios::binary);
        if (ifs.fail()) { cerr << "ERROR: given input file does not exist\n"; return 5; }
        inputs.emplace_back(istreambuf_iterator<char>(ifs), istreambuf_iterator<char>());
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
    }
    for (size_t i = 0; i < outputs.size(); i++) {
        if (status[i] != 0) {
            cerr << "ERROR: " << hc::statusMessage(status[i]) << "\n";
            return status[i];
        }
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
